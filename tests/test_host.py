"""CPU tests (-m "not gpu"): the C-ABI library loads and exports what include/lpformer_b200.h declares,
host-side graph tables match the oracle, the module mirrors the reference's state_dict, the host PPR tool is
bit-exact against the reference's numba kernel (golden PPR), and the product refuses to run without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import lpformer_oracle as O
from oracle.golden import GOLDEN_CASES, Golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lpformer_b200.build import build
    build()
    from lpformer_b200 import _lib
    return _lib.load()


def test_cabi_exports_every_declared_symbol(lib):
    from lpformer_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "lpformer_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lpf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.lpf_abi_version() == 1
    assert lib.lpf_device_ok() in (0, 1)


def test_no_cpu_fallback():
    import lpformer_b200 as L
    from lpformer_b200 import ops
    from lpformer_b200._lib import LpfError
    with pytest.raises(LpfError):
        ops.gemm(torch.zeros(4, 4), torch.zeros(4, 4))
    g = Golden("all_d32")
    model = L.LinkTransformer(g.train_args(), g.data_dict(), device="cpu").eval()
    with pytest.raises(LpfError):
        model.calc_pairwise(torch.from_numpy(g["links"]), torch.from_numpy(g["X_node"]))


def test_argument_validation(lib):
    # bad mode / NULL pointers are rejected before any launch (no GPU needed)
    rc = lib.lpf_select_count(None, 4, None, None, None, None, None, 0.0, 0.0, 0.0, 2, 0, None, None, None)
    assert rc == -1 and b"NULL" in lib.lpf_last_error()
    rc = lib.lpf_gemm(None, 4, None, 4, None, 1.0, None, 4, 4, 4, 4, 0, None)
    assert rc == -1
    assert lib.lpf_gemm(None, 4, None, 4, None, 1.0, None, 4, 0, 4, 4, 0, None) == 0      # M == 0 is a no-op
    assert lib.lpf_scan_scratch_bytes(0) == 8 and lib.lpf_scan_scratch_bytes(5000) == 24


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_state_dict_and_tables(name):
    import lpformer_b200 as L
    g = Golden(name)
    data = g.data_dict()
    model = L.LinkTransformer(g.train_args(), data, device="cpu")
    msd, ssd = g.state_dicts()
    assert set(model.state_dict()) == set(msd)                       # exact reference key set
    model.load_state_dict(msd, strict=True)
    score = L.mlp_score(model.out_dim, model.out_dim, 1, 2)
    score.load_state_dict(ssd, strict=True)
    assert model.mask == g.cfg["mask"] and model.out_dim == 2 * g.cfg["dim"]
    adj, adj_w, ppr = g.oracle_graph()
    c = L.csr_from_sparse(data["adj_mask"], mask=True)
    assert np.array_equal(c.rowptr.numpy(), adj.indptr) and np.array_equal(c.col.numpy(), adj.indices)
    p = L.csr_from_sparse(data["ppr"])
    assert np.array_equal(p.col.numpy(), ppr.indices) and np.array_equal(p.val.numpy(), ppr.val)
    an = L.gcn_normalise(L.csr_from_sparse(data["adj_t"]))
    r, cc, v = O.gcn_norm(adj_w)
    o = np.lexsort((cc, r))
    assert np.array_equal(an.col.numpy(), cc[o])
    np.testing.assert_allclose(an.val.numpy(), v[o], rtol=1e-6)


def test_csr_from_unsorted_duplicate_coo():
    import lpformer_b200 as L
    row = torch.tensor([2, 0, 2, 0, 1, 2])
    col = torch.tensor([1, 2, 1, 0, 1, 0])
    val = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    c = L.csr_from_coo(row, col, val, 3)
    assert c.rowptr.tolist() == [0, 2, 3, 5]
    assert c.col.tolist() == [0, 2, 1, 0, 1]
    assert c.val.tolist() == [4.0, 2.0, 5.0, 6.0, 4.0]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_host_ppr_tool_bit_exact_vs_reference_kernel(name, lib):
    from lpformer_b200 import synthetic as S
    g = Golden(name)
    adj, _, ppr = g.oracle_graph()
    rp, c, v = S.ppr_push(adj.indptr, adj.indices.astype(np.int32), g.cfg["alpha"], g.cfg["eps"], nthreads=3)
    assert np.array_equal(rp, ppr.indptr)
    assert np.array_equal(c, ppr.indices)
    assert np.array_equal(v.view(np.uint32), ppr.val.view(np.uint32))


def test_synthetic_graph_and_queries(lib):
    from lpformer_b200 import synthetic as S
    g = S.make_graph("citation2", scale=0.002, heldout=64)
    assert g.indptr[-1] == g.indices.size == 2 * g.edges.shape[1]
    row = np.repeat(np.arange(g.n), np.diff(g.indptr))
    assert np.all(row != g.indices)                                   # no self loops
    key = row * g.n + g.indices
    assert np.all(np.diff(key) > 0)                                   # sorted, duplicate free
    assert set(zip(row.tolist(), g.indices.tolist())) == set(zip(g.indices.tolist(), row.tolist()))  # symmetric
    held = set(zip(g.heldout[0].tolist(), g.heldout[1].tolist()))
    assert not (held & set(zip(row.tolist(), g.indices.tolist())))    # positives are not in the graph
    links = S.citation2_queries(g, 5, 7)
    assert links.shape == (2, 40)
    assert np.all(links[0].reshape(5, 8) == links[0].reshape(5, 8)[:, :1])   # one source per query
    h = S.heart_queries(g, 3, 6)
    assert h.shape == (2, 21) and h.min() >= 0 and h.max() < g.n
    # PPR of the generator equals the oracle's pure-python push on the same graph
    small = S.make_graph("cora", scale=0.05, heldout=8)
    small_ppr = O.ppr_push(small.indptr, small.indices, 0.15, 1e-4)
    rp, c, v = S.ppr_push(small.indptr, small.indices, 0.15, 1e-4)
    assert np.array_equal(c, small_ppr.indices) and np.array_equal(v.view(np.uint32), small_ppr.val.view(np.uint32))


def test_ppr_gpu_entry_points_have_no_cpu_fallback():
    """lpformer_b200.ppr (GPU PPR precompute): CPU tensors are refused, and the table-size helper follows
    2 (1 + 1 / (alpha eps)) rounded up to a power of two."""
    import torch
    from lpformer_b200 import _lib, ppr
    lib = _lib.load()
    assert lib.lpf_ppr_push_slots(0.15, 2.5e-3) == 8192          # 2 * (1 + 2666.7) = 5335.3 -> 8192
    assert lib.lpf_ppr_push_slots(0.15, 5e-5) == 1 << 19         # 2 * (1 + 133333.3) = 266668.7 -> 524288
    assert lib.lpf_ppr_push_slots(0.15, 1e-9) == -1              # beyond 2^24 keys: host tool
    assert lib.lpf_ppr_push_slots(1.5, 1e-3) == -1
    assert lib.lpf_ppr_push_scratch_bytes(8192, 8) >= 8192 * 8 * 32
    indptr = torch.tensor([0, 1, 2], dtype=torch.int64)
    indices = torch.tensor([1, 0], dtype=torch.int32)
    with pytest.raises(_lib.LpfError):
        ppr.ppr_push(indptr, indices, 0.15, 1e-3)
    ip, ix = ppr.csr_from_edge_index(torch.tensor([[1, 0, 1, 2], [0, 1, 0, 1]]), 3)       # a duplicate, unsorted
    assert ip.tolist() == [0, 1, 2, 3] and ix.tolist() == [1, 0, 1]
