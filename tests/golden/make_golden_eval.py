"""Golden vectors of the reference's EVAL DRIVERS (train/testing.py, train/evaluation.py), from the unmodified
reference run on CPU through oracle/shims/ (build container only):   python tests/golden/make_golden_eval.py

Model, graphs (train + full) and weights are those of the golden case `testset_d32`; on top of it a small dataset
dict in the layout of util/read_datasets.py: train_pos_val / valid_pos / test_pos [P, 2]; HeaRT negatives [P, K, 2]
(:150-176); plain negatives [M, 2]; citation2 negatives [P, 1000] (testing.py:21-23 hard-codes 1,000).  Stored:
the predictions of every split through test_edge / test_heart_negatives / test_edge_citation2 and the result dicts
of test(heart=False), test(heart=True), test_citation2, plus evaluate_auc on the plain split.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle", "shims"), "/root/reference/src"]

from models.link_transformer import LinkTransformer  # noqa: E402  (reference)
from models.other_models import mlp_score  # noqa: E402  (reference)
from ogb.linkproppred import Evaluator  # noqa: E402  (shim restating OGB's evaluator)
from train import testing as T  # noqa: E402  (reference)
from train.evaluation import evaluate_auc  # noqa: E402  (reference)
from oracle.golden import Golden  # noqa: E402


def main():
    g = Golden("testset_d32")
    n = g.n
    rng = np.random.default_rng(77)
    data = g.data_dict("cpu")
    model = LinkTransformer(g.train_args(), data, device="cpu").eval()
    score = mlp_score(model.out_dim, model.out_dim, 1, 2).eval()
    msd, ssd = g.state_dicts("cpu")
    model.load_state_dict(msd)
    score.load_state_dict(ssd)

    e = g["edges"]
    fe = g["full_edges"][:, e.shape[1]:]                 # the held-out (validation) edges of the full graph
    P = 24
    pick = lambda arr, k: torch.from_numpy(arr[:, rng.choice(arr.shape[1], k, replace=False)].T.copy())   # noqa: E731
    data["train_pos_val"] = pick(e, P)
    data["valid_pos"] = pick(fe, P)
    data["test_pos"] = torch.from_numpy(rng.integers(0, n, (P, 2)))
    out = {k: data[k].numpy() for k in ("train_pos_val", "valid_pos", "test_pos")}

    def heart_negs(pos, k):          # [P, K, 2]: corrupt the target (first half) or the source (second half)
        neg = np.repeat(pos.numpy()[:, None, :], k, axis=1)
        neg[:, : k // 2, 1] = rng.integers(0, n, (pos.shape[0], k // 2))
        neg[:, k // 2:, 0] = rng.integers(0, n, (pos.shape[0], k - k // 2))
        return torch.from_numpy(neg)

    res = {}
    with torch.no_grad():
        # ---- plain (non-HeaRT) protocol
        data["valid_neg"] = torch.from_numpy(rng.integers(0, n, (150, 2)))
        data["test_neg"] = torch.from_numpy(rng.integers(0, n, (150, 2)))
        out["plain_valid_neg"], out["plain_test_neg"] = data["valid_neg"].numpy(), data["test_neg"].numpy()
        ev_hit, ev_mrr = Evaluator("ogbl-collab"), Evaluator("ogbl-citation2")
        res["plain"] = T.test(model, score, data, ev_hit, ev_mrr, 64, k_list=[20, 50, 100], heart=False)
        for split, ts in (("train_pos_val", False), ("valid_pos", False), ("test_pos", True), ("valid_neg", False), ("test_neg", True)):
            out["plain_pred_" + split] = T.test_edge(model, score, data[split], 64, test_set=ts).numpy()
        pos_t, neg_t = out["plain_pred_test_pos"], out["plain_pred_test_neg"]
        res["auc"] = evaluate_auc(torch.from_numpy(np.concatenate([pos_t, neg_t])),
                                  torch.cat([torch.ones(len(pos_t)), torch.zeros(len(neg_t))]))
        # ---- HeaRT protocol
        data["valid_neg"], data["test_neg"] = heart_negs(data["valid_pos"], 40), heart_negs(data["test_pos"], 40)
        out["heart_valid_neg"], out["heart_test_neg"] = data["valid_neg"].numpy(), data["test_neg"].numpy()
        res["heart"] = T.test(model, score, data, ev_hit, ev_mrr, 200, k_list=[20, 50, 100], heart=True)
        out["heart_pred_valid_neg"] = T.test_heart_negatives(data["valid_neg"], model, score, batch_size=200).numpy()
        out["heart_pred_test_neg"] = T.test_heart_negatives(data["test_neg"], model, score, batch_size=200, test_set=True).numpy()
        # ---- citation2 protocol (1,000 negative targets per positive's source)
        c2 = {"valid_pos": data["valid_pos"][:5], "test_pos": data["test_pos"][:5], "train_pos_val": data["train_pos_val"][:5],
              "valid_neg": torch.from_numpy(rng.integers(0, n, (5, 1000))), "test_neg": torch.from_numpy(rng.integers(0, n, (5, 1000)))}
        for k, v in c2.items():
            out["c2_" + k] = v.numpy()
        res["citation2"] = T.test_citation2(model, score, c2, ev_hit, ev_mrr, 700)
        h = model.propagate()
        out["c2_pred_valid_neg"] = T.test_edge_citation2(model, score, c2["valid_pos"], h, 700, mrr_mode=True, negative_data=c2["valid_neg"]).numpy()
        out["c2_pred_test_neg"] = T.test_edge_citation2(model, score, c2["test_pos"], h, 700, mrr_mode=True, negative_data=c2["test_neg"], test=True).numpy()
        out["c2_pred_test_pos"] = T.test_edge_citation2(model, score, c2["test_pos"], h, 700, test=True).numpy()
    out["results"] = json.dumps(res)
    np.savez_compressed(os.path.join(HERE, "eval", "eval_testset_d32.npz"), **out)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "eval"), exist_ok=True)
    main()
