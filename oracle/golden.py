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

    def sets(self, prefix=""):
        """Selected sets as compute_node_mask returned them; prefix "ts_" = test_set=True (full graph tables), "am_" =
        caller-supplied adjacency (the train graph without the batch's positives)."""
        return {t: (self.z[f"{prefix}set_{t}_ix"], self.z[f"{prefix}set_{t}_src"], self.z[f"{prefix}set_{t}_tgt"])
                for t in ("cn", "1hop", "non1hop") if f"{prefix}set_{t}_ix" in self.z.files}

    @property
    def has_full_graph(self):
        return "full_edges" in self.z.files

    @property
    def has_emb(self):
        return "emb_weight" in self.z.files

    def features(self):
        """Input features of the GCN: data['x'], or data['emb'](data['x']) for the embedding case."""
        return self.z["emb_weight"][self.z["x"]] if self.has_emb else self.z["x"]

    # -- numpy CSR inputs for the oracle
    def oracle_graph(self, full=False):
        """(0/1 adjacency, weighted adjacency, PPR) as numpy CSR; full=True: the test-phase tables."""
        from . import lpformer_oracle as O
        n = self.n
        e, ew = (self["full_edges"], self["full_edge_weight"]) if full else (self["edges"], self["edge_weight"])
        row = np.concatenate([e[0], e[1]])
        col = np.concatenate([e[1], e[0]])
        w = np.concatenate([ew, ew])
        adj = O.CSR.from_coo(row, col, None, n)
        adj_w = O.CSR.from_coo(row, col, w, n)
        pre = "ppr_test_" if full else "ppr_"
        ppr = O.CSR.from_coo(self[pre + "row"], self[pre + "col"], self[pre + "val"], n)
        return adj, adj_w, ppr

    def masked_adjacency(self):
        """numpy CSR of the caller-supplied adjacency of the am_ outputs: the train graph without `am_removed`."""
        from . import lpformer_oracle as O
        e, n = self["edges"], self.n
        row = np.concatenate([e[0], e[1]])
        col = np.concatenate([e[1], e[0]])
        rm = set((self["am_removed"][0] * n + self["am_removed"][1]).tolist())
        keep = np.array([k not in rm for k in (row * n + col).tolist()])
        return O.CSR.from_coo(row[keep], col[keep], None, n)

    def masked_adjacency_coo(self, device="cpu"):
        """The same table as the coalesced 0/1 sparse COO tensor a caller passes as adj_mask=."""
        import torch
        m = self.masked_adjacency()
        row = np.repeat(np.arange(self.n), np.diff(m.indptr))
        ei = torch.from_numpy(np.stack([row, m.indices.astype(np.int64)]))
        return torch.sparse_coo_tensor(ei, torch.ones(ei.shape[1]), (self.n, self.n)).coalesce().bool().int().to(device)

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
        if self.has_full_graph:
            fe = self["full_edges"]
            fei = torch.from_numpy(np.concatenate([fe, fe[::-1]], 1))
            few = torch.from_numpy(np.concatenate([self["full_edge_weight"], self["full_edge_weight"]]))
            data["full_adj_t"] = torch.sparse_coo_tensor(fei, few, (n, n)).coalesce()
            data["full_adj_mask"] = torch.sparse_coo_tensor(fei, torch.ones(fei.shape[1]), (n, n)).coalesce().bool().int()
            data["ppr_test"] = torch.sparse_coo_tensor(
                torch.from_numpy(np.stack([self["ppr_test_row"], self["ppr_test_col"]])),
                torch.from_numpy(self["ppr_test_val"]), (n, n)).coalesce()
        data = {k: v.to(device) for k, v in data.items()}
        if self.has_emb:
            emb = torch.nn.Embedding.from_pretrained(torch.from_numpy(self["emb_weight"]), freeze=True).to(device)
            data["emb"] = emb
        return data

    def train_args(self):
        keys = ("dim", "num_heads", "trans_layers", "gnn_layers", "residual", "layer_norm", "relu",
                "thresh_cn", "thresh_1hop", "thresh_non1hop")
        return {k: self.cfg[k] for k in keys}

    def state_dicts(self, device="cpu"):
        import torch
        m = {k: torch.from_numpy(v).to(device) for k, v in self.model_params.items()}
        s = {k: torch.from_numpy(v).to(device) for k, v in self.score_params.items()}
        return m, s
