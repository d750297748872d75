"""Generate golden vectors by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

For every case below it builds a seeded degree-skewed synthetic graph, computes
PPR with the reference's own numba kernel (util/calc_ppr_scores.py:137-192 via
get_calc_ppr()), instantiates the reference LinkTransformer + mlp_score
(models/link_transformer.py, models/other_models.py) through the shims in
oracle/shims/, and stores inputs and outputs in tests/golden/<case>.npz:

  inputs : edges, edge weights, x, PPR COO, state_dicts, links, cfg (json)
  outputs: X_node (propagate), per-type selected sets (ix, src_ppr, tgt_ppr) from
           compute_node_mask, counts, elementwise/pairwise features, scores (probabilities
           and the pre-sigmoid logits), last-layer attention weights.

Cases flagged `full_graph` carry a SECOND graph (train edges + held-out validation edges, its own
PPR table) as data['full_adj_t'] / ['full_adj_mask'] / ['ppr_test'] and store the test_set=True
outputs (prefix ts_: link_transformer.py:389-406, train/testing.py:105,114) and the outputs for a
caller-supplied adjacency with the batch's positives removed (prefix am_: train_model.py:42-59,
link_transformer.py:226-254 vs :443-447).  The case flagged `emb` feeds node ids through
data['emb'] (link_transformer.py:122-123).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, "oracle", "shims"), "/root/reference/src"]

from models.link_transformer import LinkTransformer  # noqa: E402  (reference)
from models.other_models import mlp_score  # noqa: E402  (reference)
from util.calc_ppr_scores import get_calc_ppr  # noqa: E402  (reference)


def skewed_graph(n, m, seed, n_isolated=3):
    """Chung-Lu style power-law graph: undirected, simple, a few isolated nodes."""
    rng = np.random.default_rng(seed)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** -0.8
    w[-n_isolated:] = 0
    w /= w.sum()
    src = rng.choice(n, size=3 * m, p=w)
    dst = rng.choice(n, size=3 * m, p=w)
    keep = src != dst
    lo, hi = np.minimum(src, dst)[keep], np.maximum(src, dst)[keep]
    key = np.unique(lo * n + hi)
    rng.shuffle(key)
    key = np.sort(key[:m])
    return np.stack([key // n, key % n]).astype(np.int64)  # [2, E], lo < hi


def csr_from_undirected(edges, n):
    row = np.concatenate([edges[0], edges[1]])
    col = np.concatenate([edges[1], edges[0]])
    order = np.lexsort((col, row))
    row, col = row[order], col[order]
    indptr = np.zeros(n + 1, np.int64)
    np.add.at(indptr, row + 1, 1)
    return np.cumsum(indptr), col, order


def reference_ppr(indptr, indices, alpha, eps):
    nb, wt = get_calc_ppr()(indptr, indices, np.diff(indptr), alpha, eps)
    row = np.concatenate([np.full(len(c), i, np.int64) for i, c in enumerate(nb)])
    col = np.concatenate([np.asarray(c, np.int64) for c in nb])
    # create_sparse_ppr_matrix (:221-241): torch.Tensor(list) -> fp32, then sorted by (row, col)
    val = torch.Tensor(np.concatenate([np.asarray(v, np.float64) for v in wt]).tolist()).numpy()
    order = np.lexsort((col, row))
    return row[order], col[order], val[order]


def make_links(edges, n, n_links, seed):
    """Positives in the graph (both directions), random pairs, self pairs, duplicates,
    pairs touching isolated nodes."""
    rng = np.random.default_rng(seed)
    k = n_links // 3
    pos = edges[:, rng.choice(edges.shape[1], k)]
    pos[:, ::2] = pos[::-1, ::2]
    rnd = rng.integers(0, n, size=(2, n_links - k - 12))
    selfp = np.tile(rng.integers(0, n, size=4), (2, 1))
    dup = np.concatenate([pos[:, :2], pos[:, :2]], 1)
    iso = np.stack([np.array([n - 1, n - 2, 0, n - 1]), np.array([0, n - 1, n - 3, n - 1])])
    links = np.concatenate([pos, rnd, selfp, dup, iso], 1)
    return links[:, rng.permutation(links.shape[1])].astype(np.int64)


def logits_of(score, feats):
    """Pre-sigmoid output of the reference mlp_score (models/other_models.py:173-179 without the last line's sigmoid)."""
    x = feats
    for lin in score.lins[:-1]:
        x = torch.relu(lin(x))
    return score.lins[-1](x).squeeze(-1)


CASES = {
    # name: (n, m, feat, cfg, eps, n_links, weighted)
    "all_d32": (300, 1500, 24, dict(dim=32, num_heads=1, trans_layers=1, gnn_layers=2, residual=True,
                                    layer_norm=True, relu=True, thresh_cn=0, thresh_1hop=1e-2,
                                    thresh_non1hop=1e-2), 2e-3, 400, False),
    "all_lowth_h2": (300, 1500, 20, dict(dim=16, num_heads=2, trans_layers=1, gnn_layers=3, residual=False,
                                         layer_norm=True, relu=True, thresh_cn=0, thresh_1hop=1e-4,
                                         thresh_non1hop=1e-3), 1e-5, 400, True),
    "onehop_d64": (300, 2500, 64, dict(dim=64, num_heads=1, trans_layers=1, gnn_layers=3, residual=True,
                                       layer_norm=True, relu=True, thresh_cn=0, thresh_1hop=1e-2,
                                       thresh_non1hop=1), 1e-3, 400, False),
    "cora_style": (250, 500, 40, dict(dim=48, num_heads=1, trans_layers=1, gnn_layers=1, residual=False,
                                      layer_norm=False, relu=False, thresh_cn=0, thresh_1hop=1e-2,
                                      thresh_non1hop=1e-2), 1e-7, 400, False),
    "zero_1hop": (200, 900, 16, dict(dim=16, num_heads=1, trans_layers=1, gnn_layers=1, residual=False,
                                     layer_norm=True, relu=True, thresh_cn=0, thresh_1hop=0,
                                     thresh_non1hop=1e-2), 1e-4, 300, False),
    "cn_thresh": (200, 1200, 16, dict(dim=16, num_heads=1, trans_layers=1, gnn_layers=2, residual=True,
                                      layer_norm=True, relu=True, thresh_cn=1e-3, thresh_1hop=1e-3,
                                      thresh_non1hop=1e-3), 1e-5, 300, False),
    "two_layers": (200, 1000, 16, dict(dim=16, num_heads=1, trans_layers=2, gnn_layers=2, residual=True,
                                       layer_norm=True, relu=True, thresh_cn=0, thresh_1hop=1e-3,
                                       thresh_non1hop=1e-2), 1e-5, 300, False),
    "testset_d32": (300, 1500, 24, dict(dim=32, num_heads=1, trans_layers=1, gnn_layers=2, residual=True,
                                        layer_norm=True, relu=True, thresh_cn=0, thresh_1hop=1e-3,
                                        thresh_non1hop=1e-2), 1e-4, 400, True),
    "emb_d16": (200, 900, 0, dict(dim=16, num_heads=1, trans_layers=1, gnn_layers=2, residual=True,
                                  layer_norm=True, relu=True, thresh_cn=0, thresh_1hop=1e-2,
                                  thresh_non1hop=1e-2), 1e-3, 300, False),
}
FULL_GRAPH = {"testset_d32"}       # second graph for test_set=True + caller-supplied adjacency


def run_case(name, n, m, feat, cfg, eps, n_links, weighted, seed):
    torch.manual_seed(seed)
    edges = skewed_graph(n, m, seed)
    indptr, indices, _ = csr_from_undirected(edges, n)
    prow, pcol, pval = reference_ppr(indptr, indices, 0.15, eps)

    rng = np.random.default_rng(seed + 7)
    w = rng.uniform(0.5, 3.0, edges.shape[1]).astype(np.float32) if weighted else np.ones(edges.shape[1], np.float32)
    ei = torch.from_numpy(np.concatenate([edges, edges[::-1]], 1))
    ew = torch.from_numpy(np.concatenate([w, w]))
    adj_t = torch.sparse_coo_tensor(ei, ew, (n, n)).coalesce()
    adj_mask = torch.sparse_coo_tensor(ei, torch.ones(ei.shape[1]), (n, n)).coalesce().bool().int()  # read_datasets.py:95
    ppr = torch.sparse_coo_tensor(torch.from_numpy(np.stack([prow, pcol])), torch.from_numpy(pval), (n, n)).coalesce()
    extra = {}
    if feat > 0:
        x = torch.randn(n, feat)
        data = {"x": x}
    else:                       # node ids through data['emb'] (link_transformer.py:122-123)
        emb = torch.nn.Embedding(n, cfg["dim"])
        x = torch.arange(n)
        data = {"x": x, "emb": emb}
        extra["emb_weight"] = emb.weight.detach().numpy().copy()
    data.update({"adj_t": adj_t, "adj_mask": adj_mask, "ppr": ppr,
                 "full_adj_t": adj_t, "full_adj_mask": adj_mask, "ppr_test": ppr})
    if name in FULL_GRAPH:
        # the "full" graph of the test phase: train edges + held-out validation edges, its own weights and PPR table
        # (util/read_datasets.py:104-113, :126-129)
        all_e = skewed_graph(n, m + m // 5, seed + 50)
        keys_train = set((edges[0] * n + edges[1]).tolist())
        add = all_e[:, [k not in keys_train for k in (all_e[0] * n + all_e[1]).tolist()]][:, : m // 6]
        full_edges = np.concatenate([edges, add], 1)
        f_indptr, f_indices, _ = csr_from_undirected(full_edges, n)
        frow, fcol, fval = reference_ppr(f_indptr, f_indices, 0.15, eps)
        fw = np.concatenate([w, rng.uniform(0.5, 3.0, add.shape[1]).astype(np.float32)])
        fei = torch.from_numpy(np.concatenate([full_edges, full_edges[::-1]], 1))
        data["full_adj_t"] = torch.sparse_coo_tensor(fei, torch.from_numpy(np.concatenate([fw, fw])), (n, n)).coalesce()
        data["full_adj_mask"] = torch.sparse_coo_tensor(fei, torch.ones(fei.shape[1]), (n, n)).coalesce().bool().int()
        data["ppr_test"] = torch.sparse_coo_tensor(torch.from_numpy(np.stack([frow, fcol])), torch.from_numpy(fval), (n, n)).coalesce()
        extra.update(full_edges=full_edges, full_edge_weight=fw, ppr_test_row=frow, ppr_test_col=fcol, ppr_test_val=fval)

    model = LinkTransformer(dict(cfg), data, device="cpu").eval()
    score = mlp_score(model.out_dim, model.out_dim, 1, 2).eval()
    # default inits leave every LayerNorm at (1, 0) and every bias of the attention at 0:
    # perturb so the golden actually exercises them.
    with torch.no_grad():
        for k, p in list(model.named_parameters()) + list(score.named_parameters()):
            if "norm" in k or "lns" in k or k.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))

    links = torch.from_numpy(make_links(edges, n, n_links, seed + 1))
    out = {}
    with torch.no_grad():
        X = model.propagate()
        infos = model.compute_node_mask(links, False, None)
        for t, info in zip(("cn", "1hop", "non1hop"), infos):
            if info is not None:
                out[f"set_{t}_ix"] = info[0].numpy().astype(np.int64)
                out[f"set_{t}_src"] = info[1].numpy()
                out[f"set_{t}_tgt"] = info[2].numpy()
        el = model.elementwise_lin(X[links[0]] * X[links[1]])
        pw, attw = model.calc_pairwise(links, X, test_set=False, return_weights=True)
        feats = torch.cat((el, pw), dim=-1)
        feats_fwd = model(links)
        assert torch.equal(feats, feats_fwd)
        prob = score(feats)
        logit = logits_of(score, feats)
        assert torch.allclose(torch.sigmoid(logit), prob, atol=1e-7)
        if name in FULL_GRAPH:
            # ---- test_set=True: full graph for selection (and for propagate when asked); the HeaRT / citation2 eval
            # loops score test links with h = propagate() on the TRAIN graph (train/testing.py:60,105)
            extra["ts_X_node"] = model.propagate(test_set=True).numpy()
            ts_infos = model.compute_node_mask(links, True, None)
            for t, info in zip(("cn", "1hop", "non1hop"), ts_infos):
                extra[f"ts_set_{t}_ix"] = info[0].numpy().astype(np.int64)
                extra[f"ts_set_{t}_src"] = info[1].numpy()
                extra[f"ts_set_{t}_tgt"] = info[2].numpy()
            ts_pw, _ = model.calc_pairwise(links, X, test_set=True)
            ts_feats = torch.cat((el, ts_pw), dim=-1)
            extra.update(ts_pw=ts_pw.numpy(), ts_prob=score(ts_feats).numpy(), ts_logit=logits_of(score, ts_feats).numpy())
            c = model.get_structure_cnts(links, ts_infos[0], ts_infos[1], ts_infos[2], test_set=True)
            extra["ts_counts"] = torch.cat([t for t in (c[0], c[1], c[2], c[3]) if t is not None], dim=-1).numpy()
            # ---- caller-supplied adjacency: the train graph without the batch's positives (train_model.py:42-59)
            ek = adj_mask.indices()[0] * n + adj_mask.indices()[1]
            lk = torch.cat((links[0] * n + links[1], links[1] * n + links[0]))
            keep = ~torch.isin(ek, lk)
            am = torch.sparse_coo_tensor(adj_mask.indices()[:, keep], adj_mask.values()[keep], (n, n)).coalesce()
            extra["am_removed"] = adj_mask.indices()[:, ~keep].numpy().astype(np.int64)
            am_infos = model.compute_node_mask(links, False, am)
            for t, info in zip(("cn", "1hop", "non1hop"), am_infos):
                extra[f"am_set_{t}_ix"] = info[0].numpy().astype(np.int64)
                extra[f"am_set_{t}_src"] = info[1].numpy()
                extra[f"am_set_{t}_tgt"] = info[2].numpy()
            am_pw, _ = model.calc_pairwise(links, X, test_set=False, adj_mask=am)
            extra["am_pw"] = am_pw.numpy()
        # counts as the reference computes them
        if model.mask == "cn":
            counts = model.get_count(infos[0][0], links, False)
        else:
            c = model.get_structure_cnts(links, infos[0], infos[1], infos[2], test_set=False)
            counts = torch.cat([t for t in (c[0], c[1], c[2], c[3]) if t is not None], dim=-1)
    out.update(
        cfg=json.dumps(dict(cfg, eps=eps, alpha=0.15, mask=model.mask)),
        edges=edges, edge_weight=w, x=x.numpy(), ppr_row=prow, ppr_col=pcol, ppr_val=pval,
        links=links.numpy(), X_node=X.numpy(), el=el.numpy(), pw=pw.numpy(), feats=feats.numpy(),
        prob=prob.numpy(), logit=logit.numpy(), counts=counts.numpy(), att_weights=attw.numpy(), **extra,
    )
    for k, v in model.state_dict().items():
        out["model." + k] = v.numpy()
    for k, v in score.state_dict().items():
        out["score." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    sizes = {t: out[f"set_{t}_ix"].shape[1] for t in ("cn", "1hop", "non1hop") if f"set_{t}_ix" in out}
    print(f"{name}: mask={model.mask} ppr nnz/row={len(pval)/n:.1f} sets={sizes} "
          f"prob range [{prob.min():.3f},{prob.max():.3f}]")


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for i, (name, args) in enumerate(CASES.items()):
        if not only or name in only:
            run_case(name, *args, seed=100 + i)
