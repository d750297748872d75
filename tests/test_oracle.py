"""Pins oracle/lpformer_oracle.py (and oracle/ref_port.py) to the golden vectors
generated from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import lpformer_oracle as O


def graph_of(g):
    return g.oracle_graph()


def test_selection_bit_exact(golden):
    g = golden
    adj, _, ppr = graph_of(g)
    mode, sets = O.select_sets(adj, ppr, g["links"], g.cfg["thresh_cn"], g.cfg["thresh_1hop"], g.cfg["thresh_non1hop"])
    assert mode == g.cfg["mask"]
    ref = g.sets()
    assert set(ref) == set(sets)
    for t, (ix, src, tgt) in ref.items():
        li, nd, qa, qb = sets[t]
        assert np.array_equal(ix[0], li), t
        assert np.array_equal(ix[1], nd), t
        assert np.array_equal(src.view(np.uint32), qa.view(np.uint32)), t   # bit-exact fp32
        assert np.array_equal(tgt.view(np.uint32), qb.view(np.uint32)), t
    counts = O.structure_counts(sets, mode, g["links"].shape[1])
    assert np.array_equal(counts, g["counts"])


def test_propagate(golden):
    g = golden
    _, adj_w, _ = graph_of(g)
    X = O.propagate(g.features(), adj_w, g.model_params, g.cfg)
    np.testing.assert_allclose(X, g["X_node"], rtol=1e-4, atol=2e-5)
    if g.has_full_graph:        # propagate(test_set=True): the full graph's weighted adjacency
        _, full_w, _ = g.oracle_graph(full=True)
        np.testing.assert_allclose(O.propagate(g.features(), full_w, g.model_params, g.cfg), g["ts_X_node"], rtol=1e-4, atol=2e-5)


def test_features_and_scores(golden):
    g = golden
    adj, _, ppr = graph_of(g)
    feats, _, counts, alpha = O.link_features(g["links"], g["X_node"], adj, ppr, g.model_params, g.cfg)
    d = g.cfg["dim"]
    np.testing.assert_allclose(feats[:, :d], g["el"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(feats[:, d:], g["pw"], rtol=1e-4, atol=2e-5)
    logit, prob = O.mlp_score(feats, g.score_params)
    np.testing.assert_allclose(prob, g["prob"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(logit, g["logit"], rtol=1e-4, atol=1e-6)      # the pre-sigmoid values: 1e-4 relative
    # attention weights (debug output of the last layer): [2, S] = (link idx, head-mean alpha)
    aw = g["att_weights"]
    np.testing.assert_allclose(alpha.mean(1), aw[1], rtol=1e-4, atol=1e-6)


def _check_sets(ref, sets):
    assert set(ref) == set(sets)
    for t, (ix, src, tgt) in ref.items():
        li, nd, qa, qb = sets[t]
        assert np.array_equal(ix[0], li) and np.array_equal(ix[1], nd), t
        assert np.array_equal(src.view(np.uint32), qa.view(np.uint32)), t
        assert np.array_equal(tgt.view(np.uint32), qb.view(np.uint32)), t


def test_test_set_tables_and_caller_adjacency(golden):
    """test_set=True reads full_adj_mask / ppr_test (link_transformer.py:389-406) — tables that DIFFER from the train
    ones in this golden — and a caller-supplied adjacency decides CN / 1-hop only (:226-254 vs :443-447)."""
    g = golden
    if not g.has_full_graph:
        pytest.skip("case has one graph only")
    th = (g.cfg["thresh_cn"], g.cfg["thresh_1hop"], g.cfg["thresh_non1hop"])
    adj, _, ppr = g.oracle_graph()
    fadj, _, fppr = g.oracle_graph(full=True)
    assert fadj.indices.size > adj.indices.size and not np.array_equal(fppr.val, ppr.val)
    mode, ts = O.select_sets(fadj, fppr, g["links"], *th)
    _check_sets(g.sets("ts_"), ts)
    assert np.array_equal(O.structure_counts(ts, mode, g["links"].shape[1]), g["ts_counts"])
    # the train-graph sets differ from the full-graph sets: a swapped table cannot go unnoticed
    assert g.sets()["cn"][0].shape != g.sets("ts_")["cn"][0].shape or not np.array_equal(g.sets()["cn"][0], g.sets("ts_")["cn"][0])
    feats, _, _, _ = O.link_features(g["links"], g["X_node"], fadj, fppr, g.model_params, g.cfg)
    d = g.cfg["dim"]
    np.testing.assert_allclose(feats[:, d:], g["ts_pw"], rtol=1e-4, atol=2e-5)
    logit, prob = O.mlp_score(feats, g.score_params)
    np.testing.assert_allclose(prob, g["ts_prob"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(logit, g["ts_logit"], rtol=1e-4, atol=1e-6)
    # caller-supplied adjacency
    am = g.masked_adjacency()
    assert am.indices.size < adj.indices.size
    _, sets_am = O.select_sets(am, ppr, g["links"], *th, adj_far=adj)
    _check_sets(g.sets("am_"), sets_am)


def test_ppr_push_matches_reference_kernel(golden):
    g = golden
    if g.cfg["eps"] < 1e-4:
        pytest.skip("pure-Python push is slow for tiny eps; covered by the host-port / GPU tests (tests/test_gpu_ppr.py)")
    adj, _, _ = graph_of(g)
    ppr = O.ppr_push(adj.indptr, adj.indices, g.cfg["alpha"], g.cfg["eps"])
    assert np.array_equal(ppr.indices, g["ppr_col"])
    assert np.array_equal(ppr.val.view(np.uint32), g["ppr_val"].view(np.uint32))


# --------------------------------------------------------------------------- #
# oracle/ref_port.py: the torch sparse-COO restatement of the reference's algorithm
# --------------------------------------------------------------------------- #
def _port_inputs(g):
    import torch
    from oracle import ref_port as R
    adj, _, ppr = g.oracle_graph()
    A = R.coo_from_csr(adj.indptr, adj.indices, None, g.n)
    Pm = R.coo_from_csr(ppr.indptr, ppr.indices, ppr.val, g.n)
    P = {k: torch.from_numpy(v) for k, v in g.model_params.items()}
    S = {k: torch.from_numpy(v) for k, v in g.score_params.items()}
    return R, A, Pm, P, S


def test_ref_port_selection_bit_exact(golden):
    import torch
    g = golden
    R, A, Pm, _, _ = _port_inputs(g)
    sets = R.select_pairs(A, Pm, torch.from_numpy(g["links"]), g.cfg["thresh_cn"], g.cfg["thresh_1hop"],
                          g.cfg["thresh_non1hop"], g.cfg["mask"])
    ref = g.sets()
    assert set(ref) == set(sets)
    for t, (ix, src, tgt) in ref.items():
        assert np.array_equal(sets[t][0].numpy(), ix), t
        assert np.array_equal(sets[t][1].numpy().view(np.uint32), src.view(np.uint32)), t
        assert np.array_equal(sets[t][2].numpy().view(np.uint32), tgt.view(np.uint32)), t


def test_ref_port_scores(golden):
    import torch
    g = golden
    R, A, Pm, P, S = _port_inputs(g)
    prob, _ = R.score_links(torch.from_numpy(g["links"]), torch.from_numpy(g["X_node"]), A, Pm, P, S, g.cfg,
                            g.cfg["mask"])
    np.testing.assert_allclose(prob.numpy(), g["prob"], rtol=1e-4, atol=1e-6)


# --------------------------------------------------------------------------- #
# The two restatements against each other beyond the seven goldens: the set semantics of
# oracle/lpformer_oracle.py and the reference's sparse-COO algebra as ported in
# oracle/ref_port.py (models/link_transformer.py:214-319, 434-481) are independent derivations; on random
# graphs, PPR tables and thresholds they must select the same (link, node, q(src), q(tgt)) triples.
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("seed,n,m,th", [(0, 40, 120, (0.0, 1e-2, 1e-2)), (1, 60, 90, (0.0, 1e-3, 1e-1)),
                                         (2, 25, 200, (1e-2, 1e-2, 1.0)), (3, 80, 60, (0.0, 5e-2, 5e-2)),
                                         (4, 30, 30, (0.0, 1e-4, 1e-3))])
def test_restatements_agree_on_random_inputs(seed, n, m, th):
    import torch
    from oracle import ref_port as R
    rng = np.random.default_rng(seed)
    e = rng.integers(0, n, (2, m))
    e = e[:, e[0] != e[1]]
    e = np.unique(np.concatenate([e, e[::-1]], 1), axis=1)              # symmetric, simple; some nodes stay isolated
    order = np.lexsort((e[1], e[0]))
    e = e[:, order]
    indptr = np.zeros(n + 1, np.int64)
    np.add.at(indptr, e[0] + 1, 1)
    indptr = np.cumsum(indptr)
    adj = O.CSR(indptr, e[1].astype(np.int64), None, n)
    ppr = O.ppr_push(adj.indptr, adj.indices, 0.15, 1e-3)
    # links: in-graph positives, random pairs, self pairs, a duplicate, an isolated endpoint if there is one
    deg = np.diff(indptr)
    iso = np.nonzero(deg == 0)[0]
    links = np.concatenate([e[:, rng.integers(0, e.shape[1], 25)], rng.integers(0, n, (2, 40)),
                            np.tile(rng.integers(0, n, 3), (2, 1))], 1)
    links = np.concatenate([links, links[:, :2]], 1)
    if len(iso):
        links = np.concatenate([links, np.array([[iso[0]], [int(np.argmax(deg))]])], 1)
    links = links.astype(np.int64)
    mode, sets = O.select_sets(adj, ppr, links, *th)
    A = R.coo_from_csr(adj.indptr, adj.indices, None, n)
    Pm = R.coo_from_csr(ppr.indptr, ppr.indices, ppr.val, n)
    got = R.select_pairs(A, Pm, torch.from_numpy(links), *th, mode)
    assert set(got) == set(sets)
    total = 0
    for t, (li, nd, qa, qb) in sets.items():
        ix, src, tgt = got[t]
        assert np.array_equal(ix[0].numpy(), li) and np.array_equal(ix[1].numpy(), nd), (t, seed)
        assert np.array_equal(src.numpy().view(np.uint32), qa.view(np.uint32)), (t, seed)
        assert np.array_equal(tgt.numpy().view(np.uint32), qb.view(np.uint32)), (t, seed)
        total += len(li)
    assert total > 0          # the case selects something
