"""The oracle against the UNMODIFIED reference run live, beyond the committed goldens.

Only where the reference is mounted (/root/reference: this build container; the GPU box does not have it, and
nothing GPU-side reads it): the reference's LinkTransformer + mlp_score are imported through oracle/shims/ exactly as
tests/golden/make_golden.py does, run on CPU on fresh seeded cases (other seeds, sizes and thresholds than the seven
goldens), and compared with oracle/lpformer_oracle.py: selected sets bit-exact, counts exact, features and scores
within 1e-4.
"""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference/src"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")


@pytest.fixture(scope="module")
def reference():
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(REPO, "oracle", "shims"), REF]
    try:
        from models.link_transformer import LinkTransformer
        from models.other_models import mlp_score
        yield LinkTransformer, mlp_score
    finally:
        sys.path[:] = saved


CASES = [
    # seed, n, m, feat, cfg
    (11, 120, 500, 12, dict(dim=16, num_heads=1, trans_layers=1, gnn_layers=2, residual=True, layer_norm=True, relu=True,
                            thresh_cn=0, thresh_1hop=5e-3, thresh_non1hop=2e-2)),
    (12, 90, 700, 10, dict(dim=8, num_heads=2, trans_layers=1, gnn_layers=1, residual=False, layer_norm=True, relu=False,
                           thresh_cn=0, thresh_1hop=1e-3, thresh_non1hop=1)),          # 1-hop mode
    (13, 150, 400, 9, dict(dim=12, num_heads=1, trans_layers=2, gnn_layers=2, residual=False, layer_norm=False, relu=True,
                           thresh_cn=2e-3, thresh_1hop=2e-3, thresh_non1hop=5e-3)),    # thresh_cn > 0, two layers
]


@pytest.mark.parametrize("seed,n,m,feat,cfg", CASES)
def test_oracle_matches_live_reference(reference, seed, n, m, feat, cfg):
    from oracle import lpformer_oracle as O
    LinkTransformer, mlp_score = reference
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    # degree-skewed simple undirected graph with a few isolated nodes
    w = np.arange(1, n + 1, dtype=np.float64) ** -0.8
    w[-3:] = 0
    w /= w.sum()
    s, d = rng.choice(n, 3 * m, p=w), rng.choice(n, 3 * m, p=w)
    keep = s != d
    key = np.unique(np.minimum(s, d)[keep] * n + np.maximum(s, d)[keep])[:m]
    edges = np.stack([key // n, key % n]).astype(np.int64)
    row = np.concatenate([edges[0], edges[1]])
    col = np.concatenate([edges[1], edges[0]])
    order = np.lexsort((col, row))
    row, col = row[order], col[order]
    indptr = np.zeros(n + 1, np.int64)
    np.add.at(indptr, row + 1, 1)
    indptr = np.cumsum(indptr)
    adj = O.CSR(indptr, col, None, n)
    ppr = O.ppr_push(indptr, col, 0.15, 2e-3)           # (pinned to the reference's numba kernel by the goldens)
    prow = np.repeat(np.arange(n), np.diff(ppr.indptr))

    ei = torch.from_numpy(np.stack([row, col]))
    adj_t = torch.sparse_coo_tensor(ei, torch.ones(ei.shape[1]), (n, n)).coalesce()
    adj_mask = adj_t.bool().int()
    ppr_t = torch.sparse_coo_tensor(torch.from_numpy(np.stack([prow, ppr.indices])), torch.from_numpy(ppr.val), (n, n)).coalesce()
    x = torch.randn(n, feat)
    data = {"x": x, "adj_t": adj_t, "adj_mask": adj_mask, "ppr": ppr_t, "full_adj_t": adj_t,
            "full_adj_mask": adj_mask, "ppr_test": ppr_t}
    model = LinkTransformer(dict(cfg), data, device="cpu").eval()
    score = mlp_score(model.out_dim, model.out_dim, 1, 2).eval()
    with torch.no_grad():
        for k, p in list(model.named_parameters()) + list(score.named_parameters()):
            if "norm" in k or "lns" in k or k.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
    links = np.concatenate([edges[:, rng.integers(0, edges.shape[1], 40)], rng.integers(0, n, (2, 60)),
                            np.tile(rng.integers(0, n, 3), (2, 1)), np.array([[n - 1, 0], [0, n - 2]])], 1).astype(np.int64)
    tl = torch.from_numpy(links)
    with torch.no_grad():
        X = model.propagate()
        infos = model.compute_node_mask(tl, False, None)
        el = model.elementwise_lin(X[tl[0]] * X[tl[1]])
        pw, _ = model.calc_pairwise(tl, X, test_set=False, return_weights=True)
        prob = score(torch.cat((el, pw), dim=-1))

    mode, sets = O.select_sets(adj, ppr, links, cfg["thresh_cn"], cfg["thresh_1hop"], cfg["thresh_non1hop"])
    assert mode == model.mask
    for t, info in zip(("cn", "1hop", "non1hop"), infos):
        if info is None:
            assert t not in sets
            continue
        li, nd, qa, qb = sets[t]
        assert np.array_equal(info[0][0].numpy(), li) and np.array_equal(info[0][1].numpy(), nd), t
        assert np.array_equal(info[1].numpy().view(np.uint32), qa.view(np.uint32)), t
        assert np.array_equal(info[2].numpy().view(np.uint32), qb.view(np.uint32)), t
    P = {k: v.detach().numpy().astype(np.float64) for k, v in model.state_dict().items()}
    Sd = {k: v.detach().numpy().astype(np.float64) for k, v in score.state_dict().items()}
    ocfg = dict(cfg, mask=mode, alpha=0.15, eps=2e-3)
    adj_w = O.CSR(indptr, col, np.ones(len(col), np.float32), n)
    Xo = O.propagate(x.numpy(), adj_w, P, ocfg)
    np.testing.assert_allclose(Xo, X.numpy(), rtol=1e-4, atol=2e-5)
    feats, _, _, _ = O.link_features(links, X.numpy(), adj, ppr, P, ocfg)
    dd = cfg["dim"]
    np.testing.assert_allclose(feats[:, :dd], el.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(feats[:, dd:], pw.numpy(), rtol=1e-4, atol=2e-5)
    _, oprob = O.mlp_score(feats, Sd)
    np.testing.assert_allclose(oprob, prob.numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("seed,n,m,feat,cfg", CASES[:2])
def test_ref_port_matches_live_reference(reference, seed, n, m, feat, cfg):
    """oracle/ref_port.py — the port bench.py times as the CPU baseline — gives the live reference's scores."""
    from oracle import lpformer_oracle as O, ref_port as R
    LinkTransformer, mlp_score = reference
    rng = np.random.default_rng(seed + 100)
    torch.manual_seed(seed + 100)
    e = rng.integers(0, n, (2, m))
    e = e[:, e[0] != e[1]]
    e = np.unique(np.concatenate([e, e[::-1]], 1), axis=1)
    e = e[:, np.lexsort((e[1], e[0]))]
    indptr = np.zeros(n + 1, np.int64)
    np.add.at(indptr, e[0] + 1, 1)
    indptr = np.cumsum(indptr)
    ppr = O.ppr_push(indptr, e[1].astype(np.int64), 0.15, 2e-3)
    prow = np.repeat(np.arange(n), np.diff(ppr.indptr))
    ei = torch.from_numpy(e.astype(np.int64))
    adj_t = torch.sparse_coo_tensor(ei, torch.ones(ei.shape[1]), (n, n)).coalesce()
    adj_mask = adj_t.bool().int()
    ppr_t = torch.sparse_coo_tensor(torch.from_numpy(np.stack([prow, ppr.indices])), torch.from_numpy(ppr.val), (n, n)).coalesce()
    data = {"x": torch.randn(n, feat), "adj_t": adj_t, "adj_mask": adj_mask, "ppr": ppr_t, "full_adj_t": adj_t,
            "full_adj_mask": adj_mask, "ppr_test": ppr_t}
    model = LinkTransformer(dict(cfg), data, device="cpu").eval()
    score = mlp_score(model.out_dim, model.out_dim, 1, 2).eval()
    links = torch.from_numpy(np.concatenate([e[:, rng.integers(0, e.shape[1], 50)], rng.integers(0, n, (2, 80))], 1).astype(np.int64))
    with torch.no_grad():
        X = model.propagate()
        el = model.elementwise_lin(X[links[0]] * X[links[1]])
        pw, _ = model.calc_pairwise(links, X, test_set=False)
        want = score(torch.cat((el, pw), dim=-1))
    A = R.coo_from_csr(indptr, e[1].astype(np.int64), None, n)
    Pm = R.coo_from_csr(ppr.indptr, ppr.indices, ppr.val, n)
    P = {k: v.detach() for k, v in model.state_dict().items()}
    Sd = {k: v.detach() for k, v in score.state_dict().items()}
    got, _ = R.score_links(links, X, A, Pm, P, Sd, dict(cfg, mask=model.mask), model.mask)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-6)
