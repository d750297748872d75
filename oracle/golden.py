"""Loader for the golden vectors in tests/golden/*.npz (TEST INFRASTRUCTURE ONLY).

The vectors were produced by tests/golden/make_golden.py from the UNMODIFIED reference
model files; this module only reads them and rebuilds the inputs in the formats the
oracle (numpy CSR) and the product (the reference's `data` dict of torch sparse tensors)
consume.
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
GOLDEN_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    """One golden case produced by tests/golden/make_golden.py from the unmodified reference."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.z = z
        self.cfg = json.loads(str(z["cfg"]))
        self.model_params = {k[len("model."):]: z[k] for k in z.files if k.startswith("model.")}
        self.score_params = {k[len("score."):]: z[k] for k in z.files if k.startswith("score.")}
        self.n = z["x"].shape[0]

    def __getitem__(self, k):
        return self.z[k]

    def sets(self):
        return {t: (self.z[f"set_{t}_ix"], self.z[f"set_{t}_src"], self.z[f"set_{t}_tgt"])
                for t in ("cn", "1hop", "non1hop") if f"set_{t}_ix" in self.z.files}

    # -- numpy CSR inputs for the oracle
    def oracle_graph(self):
        from . import lpformer_oracle as O
        e, n = self["edges"], self.n
        row = np.concatenate([e[0], e[1]])
        col = np.concatenate([e[1], e[0]])
        w = np.concatenate([self["edge_weight"], self["edge_weight"]])
        adj = O.CSR.from_coo(row, col, None, n)
        adj_w = O.CSR.from_coo(row, col, w, n)
        ppr = O.CSR.from_coo(self["ppr_row"], self["ppr_col"], self["ppr_val"], n)
        return adj, adj_w, ppr

    # -- the reference's `data` dict (util/read_datasets.py:24-148), torch sparse COO tensors
    def data_dict(self, device="cpu"):
        import torch
        e, n = self["edges"], self.n
        ei = torch.from_numpy(np.concatenate([e, e[::-1]], 1))
        ew = torch.from_numpy(np.concatenate([self["edge_weight"], self["edge_weight"]]))
        adj_t = torch.sparse_coo_tensor(ei, ew, (n, n)).coalesce()
        adj_mask = torch.sparse_coo_tensor(ei, torch.ones(ei.shape[1]), (n, n)).coalesce().bool().int()
        ppr = torch.sparse_coo_tensor(torch.from_numpy(np.stack([self["ppr_row"], self["ppr_col"]])),
                                      torch.from_numpy(self["ppr_val"]), (n, n)).coalesce()
        x = torch.from_numpy(self["x"])
        data = {"x": x, "adj_t": adj_t, "adj_mask": adj_mask, "ppr": ppr,
                "full_adj_t": adj_t, "full_adj_mask": adj_mask, "ppr_test": ppr}
        return {k: v.to(device) for k, v in data.items()}

    def train_args(self):
        keys = ("dim", "num_heads", "trans_layers", "gnn_layers", "residual", "layer_norm", "relu",
                "thresh_cn", "thresh_1hop", "thresh_non1hop")
        return {k: self.cfg[k] for k in keys}

    def state_dicts(self, device="cpu"):
        import torch
        m = {k: torch.from_numpy(v).to(device) for k, v in self.model_params.items()}
        s = {k: torch.from_numpy(v).to(device) for k, v in self.score_params.items()}
        return m, s
