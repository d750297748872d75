"""-m gpu: the eval drivers of lpformer_b200.evaluate (same names / arguments as reference train/testing.py) against the
predictions and result dicts the UNMODIFIED reference produced for the same dataset dict
(tests/golden/make_golden_eval.py): plain protocol (test_edge), HeaRT (test_heart_negatives), citation2
(test_edge_citation2 / test_citation2)."""
import json
import os

import numpy as np
import pytest
import torch

import lpformer_b200.evaluate as E
from oracle.golden import Golden

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-4


@pytest.fixture(scope="module")
def setup():
    import lpformer_b200 as L
    g = Golden("testset_d32")
    dev = torch.device("cuda:0")
    model = L.LinkTransformer(g.train_args(), g.data_dict(dev), device=dev).to(dev).eval()
    score = L.mlp_score(model.out_dim, model.out_dim, 1, 2).to(dev).eval()
    msd, ssd = g.state_dicts(dev)
    model.load_state_dict(msd)
    score.load_state_dict(ssd)
    z = np.load(os.path.join(REPO, "tests", "golden", "eval", "eval_testset_d32.npz"))
    return model, score, z, json.loads(str(z["results"]))


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, want, **kw):
    # the drivers run our own propagate(): X_node is within 1e-4 of the reference's, and the scores follow it
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=5e-4, atol=2e-6, **kw)


def test_plain_protocol(setup):
    model, score, z, res = setup
    data = {k: T(z[k]) for k in ("train_pos_val", "valid_pos", "test_pos")}
    data["valid_neg"], data["test_neg"] = T(z["plain_valid_neg"]), T(z["plain_test_neg"])
    for split, ts in (("train_pos_val", False), ("valid_pos", False), ("test_pos", True), ("valid_neg", False), ("test_neg", True)):
        close(E.test_edge(model, score, data[split], 64, test_set=ts), z["plain_pred_" + split], err_msg=split)
    got = E.test(model, score, data, None, object(), 64, k_list=[20, 50, 100], heart=False)
    for k, want in res["plain"].items():
        np.testing.assert_allclose(got[k], want, atol=0.05, err_msg=k)        # (rank flips of near-tied scores)
    res2, hits = E.test(model, score, data, None, object(), 64, k_list=[100], heart=False, dump_test=True, metric="Hits@50")
    assert hits.shape == (data["test_pos"].shape[0],) and set(res2) == {"Hits@100", "MRR"}


def test_heart_protocol(setup):
    model, score, z, res = setup
    data = {k: T(z[k]) for k in ("train_pos_val", "valid_pos", "test_pos")}
    data["valid_neg"], data["test_neg"] = T(z["heart_valid_neg"]), T(z["heart_test_neg"])
    close(E.test_heart_negatives(data["valid_neg"], model, score, batch_size=200), z["heart_pred_valid_neg"])
    close(E.test_heart_negatives(data["test_neg"], model, score, batch_size=200, test_set=True), z["heart_pred_test_neg"])
    got = E.test(model, score, data, None, object(), 200, k_list=[20, 50, 100], heart=True)
    np.testing.assert_allclose(got["MRR"], res["heart"]["MRR"], rtol=2e-2)


def test_citation2_protocol(setup):
    model, score, z, res = setup
    c2 = {k: T(z["c2_" + k]) for k in ("valid_pos", "test_pos", "train_pos_val", "valid_neg", "test_neg")}
    h = model.propagate()
    close(E.test_edge_citation2(model, score, c2["valid_pos"], h, 700, mrr_mode=True, negative_data=c2["valid_neg"]), z["c2_pred_valid_neg"])
    close(E.test_edge_citation2(model, score, c2["test_pos"], h, 700, mrr_mode=True, negative_data=c2["test_neg"], test=True), z["c2_pred_test_neg"])
    close(E.test_edge_citation2(model, score, c2["test_pos"], h, 700, test=True), z["c2_pred_test_pos"])
    got = E.test_citation2(model, score, c2, None, object(), 700)
    np.testing.assert_allclose(got["MRR"], res["citation2"]["MRR"], rtol=2e-2)
