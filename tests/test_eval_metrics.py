"""Ranking metrics of lpformer_b200.evaluate (torch ops; CPU here) against the eval golden of the UNMODIFIED reference
(tests/golden/make_golden_eval.py ran train/testing.py + train/evaluation.py), against sklearn, and — where the
reference is mounted — against train/evaluation.py run live."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import lpformer_b200.evaluate as E

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"


@pytest.fixture(scope="module")
def ev():
    z = np.load(os.path.join(REPO, "tests", "golden", "eval", "eval_testset_d32.npz"))
    return z, json.loads(str(z["results"]))


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_plain_protocol_metrics_match_reference_results(ev):
    z, res = ev
    got = E.get_metric_score(None, object(), t(z["plain_pred_train_pos_val"]), t(z["plain_pred_valid_pos"]),
                             t(z["plain_pred_valid_neg"]), t(z["plain_pred_test_pos"]), t(z["plain_pred_test_neg"]),
                             k_list=[20, 50, 100])
    for k, want in res["plain"].items():
        np.testing.assert_allclose(got[k], want, rtol=1e-6, atol=1e-7, err_msg=k)
    pred = np.concatenate([z["plain_pred_test_pos"], z["plain_pred_test_neg"]])
    true = np.concatenate([np.ones(len(z["plain_pred_test_pos"])), np.zeros(len(z["plain_pred_test_neg"]))])
    assert E.evaluate_auc(t(pred), t(true)) == res["auc"]


def test_heart_and_citation2_metrics_match_reference_results(ev):
    z, res = ev
    # test(heart=True) scores the positives through test_edge (plain_pred_*), the negatives through test_heart_negatives
    got = E.get_metric_score_citation2(object(), t(z["plain_pred_train_pos_val"]), t(z["plain_pred_valid_pos"]),
                                       t(z["heart_pred_valid_neg"]), t(z["plain_pred_test_pos"]), t(z["heart_pred_test_neg"]))
    np.testing.assert_allclose(got["MRR"], res["heart"]["MRR"], rtol=1e-6)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_auc_ap_hits_vs_sklearn_and_definitions(seed):
    from sklearn.metrics import average_precision_score, roc_auc_score
    rng = np.random.default_rng(seed)
    n = 400
    pred = np.round(rng.random(n), 2 if seed else 6).astype(np.float32)          # seed > 0: many tied scores
    true = (rng.random(n) < 0.3).astype(np.float32)
    got = E.evaluate_auc(t(pred), t(true))
    assert got == {"AUC": round(roc_auc_score(true, pred), 4), "AP": round(average_precision_score(true, pred), 4)}
    pos, neg = t(pred[true > 0]), t(pred[true == 0])
    hits = E.evaluate_hits(None, pos, neg, [1, 10, 100, 1000])
    srt = np.sort(neg.numpy())[::-1]
    for k in (1, 10, 100):
        assert hits[f"Hits@{k}"] == pytest.approx(float((pos.numpy() > srt[k - 1]).mean()))
    assert hits["Hits@1000"] == 1.0                                               # fewer than K negatives (OGB)
    negs = t(rng.random((len(pos), 50)).astype(np.float32))
    rank = E.get_ranking_list(pos, negs).numpy()
    want = 0.5 * ((negs.numpy() >= pos.numpy()[:, None]).sum(1) + (negs.numpy() > pos.numpy()[:, None]).sum(1)) + 1
    assert np.array_equal(rank, want.astype(np.float32))
    slh = E.sample_level_hits(pos, negs)
    assert np.array_equal(slh["Hits@20"].numpy(), (want <= 20).astype(np.float32))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")
def test_metrics_vs_live_reference_evaluation_module():
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(REPO, "oracle", "shims"), REF]
    try:
        from train import evaluation as R
        from ogb.linkproppred import Evaluator
    finally:
        sys.path[:] = saved
    g = torch.Generator().manual_seed(5)
    pos, neg = torch.rand(37, generator=g), torch.rand(37, 60, generator=g)
    assert E.evaluate_mrr(pos, neg) == pytest.approx(R.evaluate_mrr(pos, neg))
    assert torch.equal(E.get_ranking_list(pos, neg), R.get_ranking_list(pos, neg))
    for k in ("Hits@20", "Hits@50", "Hits@100"):
        assert torch.equal(E.sample_level_hits(pos, neg)[k], R.sample_level_hits(pos, neg)[k])
    flat = torch.rand(500, generator=g)
    assert E.evaluate_hits(None, pos, flat, [20, 50, 100]) == pytest.approx(R.evaluate_hits(Evaluator("ogbl-collab"), pos, flat, [20, 50, 100]))
    mine = E.get_metric_score(None, object(), pos, pos * 0.9, flat[:37 * 4], pos * 1.1, flat, [20, 100])
    ref = R.get_metric_score(Evaluator("ogbl-collab"), object(), pos, pos * 0.9, flat[:37 * 4], pos * 1.1, flat, [20, 100])
    for k in ref:
        np.testing.assert_allclose(mine[k], ref[k], rtol=1e-6)
    c2 = E.get_metric_score_citation2(object(), pos, pos * 0.9, neg, pos * 1.1, neg * 1.05)
    np.testing.assert_allclose(c2["MRR"], R.get_metric_score_citation2(object(), pos, pos * 0.9, neg, pos * 1.1, neg * 1.05)["MRR"], rtol=1e-6)
    true = (torch.rand(500, generator=g) < 0.4).float()
    assert E.evaluate_auc(flat, true) == R.evaluate_auc(flat, true)
