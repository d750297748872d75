"""Stand-in for ogb.linkproppred (TEST INFRASTRUCTURE ONLY; `ogb` is an un-vendored, unpinned dependency of the
reference: requirements.txt:7).  `Evaluator` restates the published Hits@K / MRR evaluation of OGB's
ogb/linkproppred/evaluate.py (`_eval_hits`, `_eval_mrr`) so that the reference's unmodified eval drivers
(train/testing.py, train/evaluation.py) run here."""
import torch


class PygLinkPropPredDataset:  # data loading is out of the oracle's scope
    def __init__(self, *a, **k):
        raise NotImplementedError("no datasets / network in this image")


class Evaluator:
    def __init__(self, name="ogbl-collab"):
        self.name = name
        self.K = {"ogbl-collab": 50, "ogbl-ddi": 20, "ogbl-ppa": 100}.get(name)
        self.eval_metric = "mrr" if name == "ogbl-citation2" else "hits@%s" % self.K

    def eval(self, input_dict):
        pos, neg = input_dict["y_pred_pos"], input_dict["y_pred_neg"]
        if self.eval_metric == "mrr":
            # OGB _eval_mrr: optimistic / pessimistic rank average of every positive among its own negatives
            pos = pos.view(-1, 1)
            rank = 0.5 * ((neg >= pos).sum(1) + (neg > pos).sum(1)) + 1
            return {"mrr_list": 1.0 / rank.float(), "hits@10_list": (rank <= 10).float()}
        # OGB _eval_hits: a positive is a hit if it scores above the K-th highest negative
        if len(neg) < self.K:
            return {"hits@%s" % self.K: 1.0}
        kth = torch.topk(neg, self.K)[0][-1]
        return {"hits@%s" % self.K: float(torch.sum(pos > kth).cpu()) / len(pos)}
