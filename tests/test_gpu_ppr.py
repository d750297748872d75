"""PPR precompute on the GPU (lpf_ppr_push) against the host port of the reference's numba
kernel (util/calc_ppr_scores.py:137-192; csrc/ppr_push.cpp, itself pinned to the numba kernel by
tests/golden/make_golden.py): same sparsity pattern, bit-identical fp32 values."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host(g, eps):
    from lpformer_b200 import synthetic as S
    return S.ppr_push(g.indptr, g.indices, 0.15, eps)


@pytest.mark.parametrize("workload,scale,eps,fps", [("citation2", 0.02, 2.5e-3, 4096), ("citation2", 0.01, 1e-4, 4096),
                                                    ("ppa", 0.004, 5e-5, 4096), ("ddi", 0.25, 5e-6, 4096),
                                                    ("citation2", 0.02, 2.5e-3, 64), ("ppa", 0.004, 5e-5, 256)])
def test_ppr_push_bit_exact_vs_host(workload, scale, eps, fps):
    """fps = hash slots of the first pass: with 64 / 256 most sources fill their table and take the second pass."""
    from lpformer_b200 import ppr, synthetic as S
    g = S.make_graph(workload, seed=3, scale=scale, heldout=64)
    dev = torch.device("cuda:0")
    want_ptr, want_col, want_val = _host(g, eps)
    got = ppr.ppr_push(torch.from_numpy(g.indptr).to(dev), torch.from_numpy(g.indices).to(dev), 0.15, eps,
                       cap=1 << 12, first_pass_slots=fps)     # a pool that is too small at first: grown and re-run
    assert np.array_equal(got.rowptr.cpu().numpy(), want_ptr)
    assert np.array_equal(got.col.cpu().numpy(), want_col)
    assert np.array_equal(got.val.cpu().numpy().view(np.uint32), want_val.view(np.uint32))
    # every row holds its own source, values lie in (0, 1]
    assert bool(((got.val > 0) & (got.val <= 1)).all())


def test_get_ppr_matrix_mirrors_reference_entry_point():
    """get_ppr_matrix(edge_index, num_nodes, alpha, eps) — the reference's signature (util/calc_ppr_scores.py:103) —
    coalesces an unsorted edge list with duplicates, and isolated nodes get the single entry (i, i) = alpha."""
    from lpformer_b200 import ppr, synthetic as S
    g = S.make_graph("citation2", seed=5, scale=0.005, heldout=16)
    dev = torch.device("cuda:0")
    deg = np.diff(g.indptr)
    row = np.repeat(np.arange(g.n), deg)
    ei = torch.from_numpy(np.stack([row, g.indices.astype(np.int64)]))
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))
    ei = torch.cat([ei[:, perm], ei[:, :100]], 1).to(dev)          # shuffled, with duplicates
    got = ppr.get_ppr_matrix(ei, g.n, 0.15, 2.5e-3)
    want_ptr, want_col, want_val = _host(g, 2.5e-3)
    assert np.array_equal(got.rowptr.cpu().numpy(), want_ptr)
    assert np.array_equal(got.col.cpu().numpy(), want_col)
    assert np.array_equal(got.val.cpu().numpy().view(np.uint32), want_val.view(np.uint32))
    iso = np.nonzero(deg == 0)[0]
    if len(iso):
        i = int(iso[0])
        a, b = int(got.rowptr[i]), int(got.rowptr[i + 1])
        assert b - a == 1 and int(got.col[a]) == i and float(got.val[a]) == np.float32(0.15)
    coo = ppr.to_sparse_coo(got)
    assert coo.is_coalesced() and coo._nnz() == got.nnz


def test_ppr_push_feeds_the_selection():
    """The device-built table drives the model exactly like the host-built one (same selected sets)."""
    import lpformer_b200 as L
    from lpformer_b200 import ppr, synthetic as S
    g = S.make_graph("citation2", seed=8, scale=0.01, heldout=64)
    dev = torch.device("cuda:0")
    data = g.data_dict(dev)
    targs = S.train_args_of(g.cfg)
    m_host = L.LinkTransformer(targs, data, device=dev).to(dev).eval()
    data2 = dict(data)
    tbl = ppr.ppr_push(torch.from_numpy(g.indptr).to(dev), torch.from_numpy(g.indices).to(dev), 0.15, g.cfg["eps"])
    data2["ppr"] = tbl
    data2["ppr_test"] = tbl
    m_dev = L.LinkTransformer(targs, data2, device=dev).to(dev).eval()
    links = torch.from_numpy(S.citation2_queries(g, 4, 200, seed=2)).to(dev)
    a, b = m_host._select(links, False), m_dev._select(links, False)
    assert torch.equal(a.ptr, b.ptr) and torch.equal(a.node, b.node)
    assert torch.equal(a.src_ppr, b.src_ppr) and torch.equal(a.tgt_ppr, b.tgt_ppr)


def test_ppr_push_bit_exact_vs_reference_numba_golden(golden):
    """Directly against the table the reference's own numba kernel produced for the golden graphs
    (tests/golden/make_golden.py ran util/calc_ppr_scores.py:137-192): same pattern, same fp32 bits."""
    from lpformer_b200 import ppr
    adj, _, want = golden.oracle_graph()
    dev = torch.device("cuda:0")
    got = ppr.ppr_push(torch.from_numpy(adj.indptr.astype(np.int64)).to(dev),
                       torch.from_numpy(adj.indices.astype(np.int32)).to(dev), golden.cfg["alpha"], golden.cfg["eps"])
    assert np.array_equal(got.rowptr.cpu().numpy(), want.indptr)
    assert np.array_equal(got.col.cpu().numpy(), want.indices)
    assert np.array_equal(got.val.cpu().numpy().view(np.uint32), want.val.view(np.uint32))


def test_ppr_push_cora_script_eps_bit_exact_vs_host_port():
    """eps = 1e-7, the Cora line of the reference's scripts (scripts/replicate_heart.sh:4): the push touches nearly every
    node of the component many times (~10^5 pops per source) — same table, bit for bit, as the host port of the
    reference's numba kernel.  (Full Cora shape, 2,708 nodes / 5.66 M entries: 19 s on a B200 vs 16 s for the host port
    on 16 cores and 447-552 s for the numba kernel on 8; `tools/ppr_cora.py`.)"""
    from lpformer_b200 import ppr, synthetic as S
    g = S.make_graph("cora", seed=3, scale=0.12, heldout=32)
    assert g.cfg["eps"] == 1e-7
    dev = torch.device("cuda:0")
    got = ppr.ppr_push(torch.from_numpy(g.indptr).to(dev), torch.from_numpy(g.indices).to(dev), 0.15, g.cfg["eps"])
    assert np.array_equal(got.rowptr.cpu().numpy(), g.ppr[0])
    assert np.array_equal(got.col.cpu().numpy(), g.ppr[1])
    assert np.array_equal(got.val.cpu().numpy().view(np.uint32), g.ppr[2].view(np.uint32))
