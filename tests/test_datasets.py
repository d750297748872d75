"""Data ingestion (lpformer_b200/datasets.py) — SURVEY 8(f) rank 4, reference util/read_datasets.py:20-254.

Three layers:
  * file formats and dict construction against plain numpy expectations on small fixture directories written by this
    file (Planetoid text files, the OGB raw download layout, HeaRT negatives, PPR cache files) — runs everywhere;
  * the same fixtures through the UNMODIFIED reference functions `read_data_planetoid` / `read_data_ogb` (imported
    from /root/reference through oracle/shims/, data directories and the PPR cache redirected by attribute, its numba
    push computing the tables) — every key of the two dicts compared; only where the reference is mounted;
  * (gpu) ingestion onto the device: the GPU push gives the host tool's table bit for bit and the model built from the
    dict selects the oracle's sets.
"""
import gzip
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from lpformer_b200 import datasets as D
from lpformer_b200.graph import CSR

REF = "/root/reference/src"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")


# ---------------------------------------------------------------------------------------------------------------
# fixture writers
# ---------------------------------------------------------------------------------------------------------------
def _simple_edges(rng, n, m):
    s, d = rng.integers(0, n, 4 * m), rng.integers(0, n, 4 * m)
    keep = s != d
    key = np.unique(np.minimum(s, d)[keep] * n + np.maximum(s, d)[keep])
    key = rng.permutation(key)[:m]
    e = np.stack([key // n, key % n]).astype(np.int64)
    flip = rng.random(e.shape[1]) < 0.5
    e[:, flip] = e[::-1, flip]
    return e


def write_planetoid(root, name="cora", n=60, m=150, feat=7, seed=0, heart=True):
    rng = np.random.default_rng(seed)
    e = _simple_edges(rng, n, m)
    e[:, 0] = (0, n - 1)                      # every id of 0..n-1 appears at least ... see below
    ids = rng.permutation(n)
    cover = np.stack([ids, np.roll(ids, 1)])  # a ring through every node, so that len(node_set) == n
    e = np.concatenate([cover, e], 1)
    key = np.unique(np.minimum(e[0], e[1]) * n + np.maximum(e[0], e[1]), return_index=True)[1]
    e = e[:, np.sort(key)]
    m = e.shape[1]
    cut1, cut2 = int(0.8 * m), int(0.9 * m)
    splits = {"train": e[:, :cut1], "valid": e[:, cut1:cut2], "test": e[:, cut2:]}
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    for s, arr in splits.items():
        with open(os.path.join(d, f"{s}_pos.txt"), "w") as fh:
            for a, b in arr.T:
                fh.write(f"{a}\t{b}\n")
            if s == "train":
                fh.write("3\t3\n")            # a self loop: counted as a node, skipped as an edge (:160-163)
    for s in ("valid", "test"):
        neg = rng.integers(0, n, (splits[s].shape[1], 2))
        with open(os.path.join(d, f"{s}_neg.txt"), "w") as fh:
            for a, b in neg:
                fh.write(f"{a}\t{b}\n")
    torch.save({"entity_embedding": torch.from_numpy(rng.standard_normal((n, feat)).astype(np.float32))},
               os.path.join(d, "gnn_feature"))
    if heart:
        hd = os.path.join(root, "heart", name)
        os.makedirs(hd, exist_ok=True)
        for s in ("valid", "test"):
            np.save(os.path.join(hd, f"heart_{s}_samples.npy"), rng.integers(0, n, (splits[s].shape[1], 5, 2)))
    return n, splits


def _gz(path, arr, fmt):
    with gzip.open(path, "wt") as fh:
        np.savetxt(fh, arr, fmt=fmt, delimiter=",")


def make_ogb_arrays(name, seed=0, n=80, m=260):
    """In-memory raw content of a small ogbl-* graph: what ogb ships on disk, before its own loader runs."""
    rng = np.random.default_rng(seed)
    if name == "ogbl-citation2":
        e = _simple_edges(rng, n, m)
        back = e[::-1, :10]                                   # ten mutual citations
        e = np.concatenate([e, back], 1)
    else:
        e = _simple_edges(rng, n, m)
        e = np.stack([e.min(0), e.max(0)])
    raw = {"edge": e, "num_nodes": n}
    if name != "ogbl-ddi":
        raw["node_feat"] = (rng.integers(0, 2, (n, 6)) if name == "ogbl-ppa" else rng.standard_normal((n, 6))).astype(np.float32)
    split = {}
    if name == "ogbl-collab":
        dup = e[:, :40]                                       # collaborations repeated in another year
        e = np.concatenate([e, dup], 1)
        raw["edge"] = e
        raw["edge_weight"] = rng.integers(1, 4, (e.shape[1], 1)).astype(np.float32)
        raw["edge_year"] = rng.integers(2000, 2016, (e.shape[1], 1)).astype(np.int64)
        split["train"] = {"edge": e.T.copy(), "weight": raw["edge_weight"][:, 0].astype(np.int64),
                          "year": raw["edge_year"][:, 0].copy()}
    elif name == "ogbl-citation2":
        split["train"] = {"source_node": e[0].copy(), "target_node": e[1].copy()}
    else:
        split["train"] = {"edge": e.T.copy()}
    q = 12
    for s in ("valid", "test"):
        if name == "ogbl-citation2":
            split[s] = {"source_node": rng.integers(0, n, q), "target_node": rng.integers(0, n, q),
                        "target_node_neg": rng.integers(0, n, (q, 9))}
        else:
            split[s] = {"edge": _simple_edges(rng, n, q).T.copy(), "edge_neg": rng.integers(0, n, (3 * q, 2))}
            if name == "ogbl-collab":
                split[s]["weight"] = np.ones(q, np.int64)
                split[s]["year"] = np.full(q, 2018 if s == "valid" else 2019, np.int64)
    return raw, split


def write_ogb_raw(root, name, raw, split, heart=True, seed=0):
    base = os.path.join(root, name.replace("-", "_"))
    os.makedirs(os.path.join(base, "raw"), exist_ok=True)
    _gz(os.path.join(base, "raw", "edge.csv.gz"), raw["edge"].T, "%d")
    _gz(os.path.join(base, "raw", "num-node-list.csv.gz"), np.array([[raw["num_nodes"]]]), "%d")
    _gz(os.path.join(base, "raw", "num-edge-list.csv.gz"), np.array([[raw["edge"].shape[1]]]), "%d")
    if "node_feat" in raw:
        _gz(os.path.join(base, "raw", "node-feat.csv.gz"), raw["node_feat"], "%.9g")
    for k in ("edge_weight", "edge_year"):
        if k in raw:
            _gz(os.path.join(base, "raw", k + ".csv.gz"), raw[k], "%d")
    sd = os.path.join(base, "split", D.OGB_META[name]["split"])
    os.makedirs(sd, exist_ok=True)
    for s, d in split.items():
        torch.save(d, os.path.join(sd, s + ".pt"))
    if heart:
        rng = np.random.default_rng(seed + 7)
        hd = os.path.join(root, "heart", name)
        os.makedirs(hd, exist_ok=True)
        for s in ("valid", "test"):
            npos = len(split[s]["source_node"] if name == "ogbl-citation2" else split[s]["edge"])
            if name == "ogbl-ppa":
                ix = torch.from_numpy(np.sort(rng.permutation(npos)[: npos // 2]))
                torch.save(ix, os.path.join(hd, f"{s}_samples_index.pt"))
                npos = npos // 2
            np.save(os.path.join(hd, f"heart_{s}_samples.npy"), rng.integers(0, raw["num_nodes"], (npos, 5, 2)))


def _args(name, **kw):
    return SimpleNamespace(**dict(dict(data_name=name, heart=False, use_val_in_test=False, eps=1e-3, dim=8), **kw))


def _csr_np(c: CSR):
    rows = np.repeat(np.arange(c.n), np.diff(c.rowptr.cpu().numpy()))
    return rows, c.col.cpu().numpy().astype(np.int64), None if c.val is None else c.val.cpu().numpy()


def _coo_np(t):
    """(row, col, val) of what the reference's dict holds: torch sparse tensor (coalesced here) or shim SparseTensor
    (duplicates summed, the form the GCN sees them in)."""
    if isinstance(t, torch.Tensor):
        t = t.coalesce()
        return t.indices()[0].numpy(), t.indices()[1].numpy(), t.values().numpy()
    r, c, v = t.coalesce().coo()
    return r.numpy(), c.numpy(), None if v is None else v.numpy()


# ---------------------------------------------------------------------------------------------------------------
# formats and dict construction (no reference needed)
# ---------------------------------------------------------------------------------------------------------------
def test_planetoid_files_to_dict(tmp_path):
    n, splits = write_planetoid(str(tmp_path))
    torch.manual_seed(3)
    data = D.read_data_planetoid(_args("cora"), "cpu", data_dir=str(tmp_path), ppr_cache_dir=str(tmp_path / "ppr"))
    assert data["num_nodes"] == n and data["dataset"] == "cora"
    tr = splits["train"]
    assert np.array_equal(data["train_pos"].numpy(), tr.T)                      # the self loop is not an edge
    assert np.array_equal(data["edge_index"].numpy(), np.concatenate([tr, tr[::-1]], 1))
    assert np.array_equal(data["valid_pos"].numpy(), splits["valid"].T)
    assert data["train_pos_val"].shape == data["valid_pos"].shape
    r, c, v = _csr_np(data["adj_t"])
    want = np.unique(np.concatenate([tr[0] * n + tr[1], tr[1] * n + tr[0]]))
    assert np.array_equal(r * n + c, want) and np.all(v == 1)
    assert data["adj_mask"].val is None and np.array_equal(data["adj_mask"].col.numpy(), c.astype(np.int32))
    assert data["full_adj_mask"] is data["adj_mask"] and data["ppr_test"] is data["ppr"]
    assert np.array_equal(data["degree"].numpy(), np.bincount(np.concatenate([tr[0], tr[1]]), minlength=n))
    # the PPR table: rows sum to < 1, the source itself holds at least alpha, sorted columns
    pr, pc, pv = _csr_np(data["ppr"])
    assert np.all(np.diff(pr * n + pc) > 0)
    diag = pv[pr == pc]
    assert len(diag) == n and np.all(diag >= np.float32(0.15))
    # the cache file has the reference's name and is what the second call loads
    path = D.ppr_cache_path("cora", 0.15, 1e-3, False, str(tmp_path / "ppr"))
    assert path.endswith(os.path.join("cora", "sparse_adj-015_eps-0001.pt")) and os.path.isfile(path)
    again = D.get_ppr("cora", torch.zeros(2, 0, dtype=torch.int64), n, 0.15, 1e-3, False, "cpu", str(tmp_path / "ppr"))
    assert torch.equal(again.col, data["ppr"].col) and torch.equal(again.val, data["ppr"].val)
    # HeaRT negatives replace the files' own
    h = D.read_data_planetoid(_args("cora", heart=True), "cpu", data_dir=str(tmp_path), ppr_cache_dir=str(tmp_path / "ppr"))
    assert h["valid_neg"].shape == (splits["valid"].shape[1], 5, 2)


def test_ppr_cache_written_by_the_reference_loads(tmp_path):
    """util/calc_ppr_scores.py:266 saves a pickled torch_sparse.SparseTensor; without torch_sparse the classes are
    stubbed while unpickling and the storage fields read."""
    import pickle
    import types
    n = 5
    row = torch.tensor([0, 0, 1, 3, 4]); col = torch.tensor([0, 2, 1, 3, 0]); val = torch.tensor([.5, .1, .9, .7, .2])
    mod_t, mod_s = types.ModuleType("torch_sparse.tensor"), types.ModuleType("torch_sparse.storage")
    pkg = types.ModuleType("torch_sparse")

    class SparseStorage:
        pass

    class SparseTensor:
        pass

    SparseStorage.__module__, SparseTensor.__module__ = "torch_sparse.storage", "torch_sparse.tensor"
    SparseStorage.__qualname__, SparseTensor.__qualname__ = "SparseStorage", "SparseTensor"
    mod_s.SparseStorage, mod_t.SparseTensor = SparseStorage, SparseTensor
    st = SparseStorage(); st._row, st._rowptr, st._col, st._value, st._sparse_sizes = row, None, col, val, (n, n)
    sp = SparseTensor(); sp.storage = st
    path = D.ppr_cache_path("toy", 0.15, 5e-5, True, str(tmp_path))
    os.makedirs(os.path.dirname(path))
    sys.modules.update({"torch_sparse": pkg, "torch_sparse.tensor": mod_t, "torch_sparse.storage": mod_s})
    try:
        torch.save(sp, path)
    finally:
        for k in ("torch_sparse", "torch_sparse.tensor", "torch_sparse.storage"):
            sys.modules.pop(k, None)
    assert path.endswith("sparse_adj-015_eps-5e-05_val.pt")
    t = D.get_ppr("toy", torch.zeros(2, 0, dtype=torch.int64), n, 0.15, 5e-5, True, "cpu", str(tmp_path))
    assert t.rowptr.tolist() == [0, 2, 3, 3, 4, 5] and t.col.tolist() == [0, 2, 1, 3, 0]
    assert torch.equal(t.val, val)


@pytest.mark.parametrize("name", ["ogbl-collab", "ogbl-ddi", "ogbl-ppa", "ogbl-citation2"])
def test_ogb_raw_layout(tmp_path, name):
    raw, split = make_ogb_arrays(name)
    write_ogb_raw(str(tmp_path), name, raw, split)
    data, se = D.read_ogb_raw(name, str(tmp_path))
    e = raw["edge"]
    assert data.num_nodes == raw["num_nodes"]
    if name == "ogbl-citation2":
        assert np.array_equal(data.edge_index.numpy(), e)
    else:                                   # every edge followed by its reverse
        assert np.array_equal(data.edge_index.numpy()[:, 0::2], e) and np.array_equal(data.edge_index.numpy()[:, 1::2], e[::-1])
    if name == "ogbl-collab":
        assert np.array_equal(data.edge_weight.numpy()[0::2], raw["edge_weight"])
        assert np.array_equal(data.edge_year.numpy()[1::2], raw["edge_year"])
    if name == "ogbl-ddi":
        assert data.x is None
    else:
        np.testing.assert_array_equal(data.x.numpy(), raw["node_feat"])
    for s in split:
        for k, v in split[s].items():
            assert np.array_equal(se[s][k].numpy(), v)


def test_ogb_dict_collab_use_val_in_test(tmp_path):
    name = "ogbl-collab"
    raw, split = make_ogb_arrays(name)
    write_ogb_raw(str(tmp_path), name, raw, split)
    torch.manual_seed(0)
    d = D.read_data_ogb(_args(name, use_val_in_test=True, heart=True), "cpu", data_dir=str(tmp_path), ppr_cache=False)
    n = raw["num_nodes"]
    keep = raw["edge_year"][:, 0] >= 2007
    e, w = raw["edge"][:, keep], raw["edge_weight"][keep, 0]
    dense = np.zeros((n, n))
    np.add.at(dense, (e[0], e[1]), w)
    np.add.at(dense, (e[1], e[0]), w)
    r, c, v = _csr_np(d["adj_t"])
    got = np.zeros((n, n)); got[r, c] = v
    assert np.array_equal(got, dense)
    assert np.array_equal(d["train_pos"].numpy(), raw["edge"].T[keep])
    ve = split["valid"]["edge"].T
    full = dense.copy()
    full[ve[0], ve[1]] += 1; full[ve[1], ve[0]] += 1
    r, c, v = _csr_np(d["full_adj_t"])
    got = np.zeros((n, n)); got[r, c] = v
    assert np.array_equal(got, full)
    r, c, _ = _csr_np(d["full_adj_mask"])
    assert np.array_equal(np.stack([r, c]), np.stack(np.nonzero(full)))
    assert d["ppr_test"] is not d["ppr"] and d["ppr_test"].nnz != d["ppr"].nnz
    assert d["degree_test"].sum() == d["degree"].sum() + 2 * ve.shape[1]
    assert d["valid_neg"].shape == (ve.shape[1], 5, 2)


# ---------------------------------------------------------------------------------------------------------------
# against the unmodified reference
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_read():
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(REPO, "oracle", "shims"), REF]
    try:
        from util import calc_ppr_scores as C, read_datasets as R

        def get_ppr(dataset, edge_index, num_nodes, alpha, eps, is_val):
            # util/calc_ppr_scores.py:244-270 without its file cache (the reference tree is read-only)
            nb, w = C.get_ppr_matrix(edge_index, num_nodes, alpha, eps)
            return C.create_sparse_ppr_matrix(nb, w).to_torch_sparse_coo_tensor()

        R.get_ppr = get_ppr
        yield R
    finally:
        sys.path[:] = saved


def _compare_dicts(ref, ours, n):
    assert set(k for k in ref if k != "full_edge_index") <= set(ours), set(ref) - set(ours)
    for k, rv in ref.items():
        ov = ours[k]
        if k in ("adj_t", "full_adj_t", "adj_mask", "full_adj_mask", "ppr", "ppr_test"):
            rr, rc, rval = _coo_np(rv)
            if "mask" in k:
                keep = rval != 0
                rr, rc = rr[keep], rc[keep]
            o_r, o_c, o_v = _csr_np(ov)
            assert np.array_equal(rr * n + rc, o_r * n + o_c), k
            if "mask" not in k:
                if "ppr" in k:      # the numba kernel's table, bit for bit
                    assert np.array_equal(rval.astype(np.float32).view(np.uint32), o_v.view(np.uint32)), k
                else:
                    assert np.array_equal(rval.astype(np.float32), o_v), k
        elif isinstance(rv, torch.Tensor):
            assert rv.shape == ov.shape and rv.dtype == ov.dtype, (k, rv.shape, ov.shape, rv.dtype, ov.dtype)
            assert torch.equal(rv.detach().cpu(), ov.detach().cpu()), k
        else:
            assert rv == ov, k


@needs_reference
@pytest.mark.parametrize("heart", [False, True])
def test_planetoid_vs_live_reference(ref_read, tmp_path, heart):
    n, _ = write_planetoid(str(tmp_path), seed=5)
    ref_read.DATA_DIR, ref_read.HEART_DIR = str(tmp_path), str(tmp_path / "heart")
    args = _args("cora", heart=heart, eps=2e-4)
    torch.manual_seed(1)
    ref = ref_read.read_data_planetoid(args, "cpu")
    torch.manual_seed(1)
    ours = D.read_data_planetoid(args, "cpu", data_dir=str(tmp_path), ppr_cache=False)
    _compare_dicts(ref, ours, n)


class _FakePygDataset:
    """What `PygLinkPropPredDataset(name)[0]` / `.get_edge_split()` hand the reference, built from the in-memory raw
    arrays by ogb's documented rules (inverse edges interleaved) — NOT through lpformer_b200's reader."""
    raw = split = meta = None

    def __init__(self, name):
        raw, meta = self.raw, self.meta
        e = raw["edge"]
        extras = {k: raw[k] for k in ("edge_weight", "edge_year") if k in raw}
        if meta["add_inverse_edge"]:
            e = np.stack([e, e[::-1]], axis=2).reshape(2, -1)
            extras = {k: np.repeat(v, 2, axis=0) for k, v in extras.items()}
        outer = self

        class Data(dict):
            def to(self, device):
                return self

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v

        d = Data(num_nodes=raw["num_nodes"], edge_index=torch.from_numpy(e.copy()))
        if "node_feat" in raw:
            d["x"] = torch.from_numpy(raw["node_feat"].copy())
        for k, v in extras.items():
            d[k] = torch.from_numpy(v.copy())
        self._data = d

    def __getitem__(self, i):
        return self._data

    def get_edge_split(self):
        return {s: {k: torch.from_numpy(np.array(v)) for k, v in d.items()} for s, d in self.split.items()}


@needs_reference
@pytest.mark.parametrize("name,heart,use_val", [("ogbl-collab", True, True), ("ogbl-collab", False, False),
                                                 ("ogbl-ddi", True, False), ("ogbl-ppa", True, False),
                                                 ("ogbl-citation2", False, False)])
def test_ogb_vs_live_reference(ref_read, tmp_path, name, heart, use_val):
    raw, split = make_ogb_arrays(name, seed=3)
    write_ogb_raw(str(tmp_path), name, raw, split)
    _FakePygDataset.raw, _FakePygDataset.split, _FakePygDataset.meta = raw, split, D.OGB_META[name]
    ref_read.PygLinkPropPredDataset = _FakePygDataset
    ref_read.DATA_DIR, ref_read.HEART_DIR = str(tmp_path), str(tmp_path / "heart")
    args = _args(name, heart=heart, use_val_in_test=use_val, eps=5e-4)
    torch.manual_seed(2)
    ref = ref_read.read_data_ogb(args, "cpu")
    torch.manual_seed(2)
    ours = D.read_data_ogb(args, "cpu", data_dir=str(tmp_path), ppr_cache=False)
    _compare_dicts(ref, ours, raw["num_nodes"])


# ---------------------------------------------------------------------------------------------------------------
# on the device
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_ingest_on_device_and_select(tmp_path):
    import lpformer_b200 as L
    from oracle import lpformer_oracle as O
    n, splits = write_planetoid(str(tmp_path), n=300, m=1200, feat=16, seed=9)
    args = _args("cora", eps=1e-4)
    host = D.read_data_planetoid(args, "cpu", data_dir=str(tmp_path), ppr_cache=False)
    dev = D.read_data_planetoid(args, "cuda", data_dir=str(tmp_path), ppr_cache=False)
    assert dev["ppr"].col.is_cuda
    for f in ("rowptr", "col"):
        assert torch.equal(getattr(dev["ppr"], f).cpu(), getattr(host["ppr"], f))
    assert torch.equal(dev["ppr"].val.cpu().view(torch.int32), host["ppr"].val.view(torch.int32))   # GPU push == host push
    cfg = dict(dim=32, num_heads=1, trans_layers=1, gnn_layers=2, residual=True, layer_norm=True, relu=True,
               thresh_cn=0, thresh_1hop=1e-3, thresh_non1hop=1e-2)
    targs = dict(cfg, mask_input=False, feat_drop=0.0, pred_dropout=0.0, gnn_drop=0.0, att_drop=0.0, dropout=0.0)
    model = L.LinkTransformer(targs, dev, device=torch.device("cuda")).cuda().eval()
    links = torch.cat([dev["test_pos"].t(), dev["test_neg"].t()], 1)
    infos = model.compute_node_mask(links, False, None)
    adj = O.CSR(host["adj_mask"].rowptr.numpy(), host["adj_mask"].col.numpy().astype(np.int64), None, n)
    ppr = O.CSR(host["ppr"].rowptr.numpy(), host["ppr"].col.numpy().astype(np.int64), host["ppr"].val.numpy(), n)
    mode, sets = O.select_sets(adj, ppr, links.cpu().numpy(), 0, 1e-3, 1e-2)
    for t, info in zip(("cn", "1hop", "non1hop"), infos):
        li, nd, qa, qb = sets[t]
        assert np.array_equal(info[0][0].cpu().numpy(), li) and np.array_equal(info[0][1].cpu().numpy(), nd), t
