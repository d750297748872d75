"""Data ingestion: the reference's `data` dict builders (util/read_datasets.py:20-254).

`read_data_planetoid(args, device)` and `read_data_ogb(args, device)` keep the reference's names, arguments
(`args.data_name`, `args.heart`, `args.use_val_in_test`, `args.eps`, `args.dim`) and the keys of the dict they return
(`train_pos`, `valid_pos`, `valid_neg`, `test_pos`, `test_neg`, `train_pos_val`, `x`, `adj_t`, `adj_mask`, `full_adj_t`,
`full_adj_mask`, `full_edge_index`, `degree`, `degree_test`, `ppr`, `ppr_test`, ...), so `train/testing.py`'s drivers
(`lpformer_b200.evaluate`) and `LinkTransformer(train_args, data, device)` take the result as they take the
reference's.  What differs is the FORM of the sparse tables: the reference keeps `adj_t` as a
torch_sparse.SparseTensor and `adj_mask` / `ppr` as N x N sparse COO tensors which every batch slices with an
O(nnz) `index_select`; here each is built once as the sorted CSR the kernels address directly
(`lpformer_b200.graph.CSR`, which `LinkTransformer` accepts wherever the reference passes a sparse tensor).

Neither `ogb` nor `torch_geometric` exists in this image, so the OGB graphs are read from the package's RAW download
layout (`<root>/ogbl_<name>/raw/*.csv.gz`, `split/<type>/{train,valid,test}.pt` — what ogb's `read_csv_graph_raw`
parses, restated in `read_ogb_raw`), not from PyG's pickled `processed/` files.

PPR tables: `get_ppr` follows util/calc_ppr_scores.py:244-270 — same cache file names under
`node_subsets/ppr/<dataset>/`, computed when absent — with the push on the GPU (`lpformer_b200.ppr`, bit-identical to
the reference's numba kernel) when the target device is CUDA and by the multi-threaded host tool of the same library
otherwise (ingestion is one-time set-up on either).
"""
from __future__ import annotations

import io
import os
import pickle
import warnings
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from .graph import CSR, csr_from_coo

DATA_DIR = os.environ.get("LPF_DATA_DIR", os.path.join(os.getcwd(), "dataset"))
PPR_DIR = os.environ.get("LPF_PPR_DIR", os.path.join(os.getcwd(), "node_subsets", "ppr"))

# ogb's master.csv rows for the four link-prediction graphs the reference's scripts use (scripts/replicate_*.sh)
OGB_META = {
    "ogbl-collab": dict(add_inverse_edge=True, split="time", edge_files=("edge_weight", "edge_year"), node_files=()),
    "ogbl-ddi": dict(add_inverse_edge=True, split="target", edge_files=(), node_files=()),
    "ogbl-ppa": dict(add_inverse_edge=True, split="throughput", edge_files=(), node_files=()),
    "ogbl-citation2": dict(add_inverse_edge=False, split="time", edge_files=(), node_files=("node_year",)),
}


# ------------------------------------------------------------------------------------------------------------------
# small tensor helpers (torch_geometric.utils.{degree, to_undirected, coalesce} as the reference uses them)
# ------------------------------------------------------------------------------------------------------------------
def degree(index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """`torch_geometric.utils.degree`: occurrences of every node id, fp32 (util/read_datasets.py:113-115,226)."""
    return torch.bincount(index.to(torch.int64), minlength=num_nodes).to(torch.float32)


def to_undirected(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor] = None, num_nodes: Optional[int] = None):
    """`torch_geometric.utils.to_undirected(edge_index[, edge_attr], reduce='add')`: both directions of every edge,
    sorted by (row, col), duplicates merged (their attributes summed)."""
    ei = edge_index.to(torch.int64)
    n = int(ei.max()) + 1 if num_nodes is None and ei.numel() else int(num_nodes or 0)
    key = torch.cat([ei[0] * n + ei[1], ei[1] * n + ei[0]])
    if edge_attr is None:
        key = torch.unique(key)
        return torch.stack([key // n, key % n])
    uniq, inv = torch.unique(key, return_inverse=True)
    attr = torch.cat([edge_attr, edge_attr])
    out = torch.zeros((uniq.numel(),) + tuple(attr.shape[1:]), dtype=attr.dtype, device=attr.device).index_add_(0, inv, attr)
    return torch.stack([uniq // n, uniq % n]), out


def _csr_pair(edge_index: torch.Tensor, weight: Optional[torch.Tensor], n: int, device):
    """(adj_t, adj_mask) of an edge list: weighted CSR (duplicate entries summed, which is what the GCN's
    normalisation and SpMM make of a SparseTensor holding them twice) and the 0/1 CSR
    (`.coalesce().bool().int()`, util/read_datasets.py:95)."""
    ei = edge_index.to(device=device, dtype=torch.int64)
    w = torch.ones(ei.size(1), dtype=torch.float32, device=ei.device) if weight is None else weight.to(ei.device).float().view(-1)
    adj_t = csr_from_coo(ei[0], ei[1], w, n)
    nz = adj_t.val != 0 if adj_t.val is not None else None
    if nz is not None and not bool(nz.all()):
        rows = torch.repeat_interleave(torch.arange(n, device=ei.device), adj_t.rowptr[1:] - adj_t.rowptr[:-1])
        mask = csr_from_coo(rows[nz], adj_t.col[nz].to(torch.int64), None, n)
    else:
        mask = CSR(adj_t.rowptr, adj_t.col, None, n)
    return adj_t, mask


# ------------------------------------------------------------------------------------------------------------------
# PPR tables with the reference's cache (util/calc_ppr_scores.py:196-270)
# ------------------------------------------------------------------------------------------------------------------
def ppr_cache_path(dataset: str, alpha: float, eps: float, is_val: bool, root: Optional[str] = None) -> str:
    """util/calc_ppr_scores.py:252-257: `node_subsets/ppr/<dataset>/sparse_adj-015_eps-5e-05[_val].pt`."""
    alpha_str = str(alpha).replace(".", "")
    eps_str = str(eps).replace(".", "")
    return os.path.join(root or PPR_DIR, dataset, f"sparse_adj-{alpha_str}_eps-{eps_str}" + ("_val" if is_val else "") + ".pt")


class _Stub:
    """Stands in for any class of a package that is not installed while unpickling (torch_sparse's SparseTensor /
    SparseStorage): keeps whatever state the pickle carries."""

    def __init__(self, *a, **k):
        self._args = a

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "torch_sparse":
            return type(name, (_Stub,), {})
        return super().find_class(module, name)


class _TolerantPickle:
    """`pickle_module` for torch.load: torch_sparse classes become stubs."""
    __name__ = "lpformer_b200_tolerant_pickle"
    Unpickler = _TolerantUnpickler
    load = staticmethod(lambda f, **k: _TolerantUnpickler(f, **k).load())


def _csr_from_reference_file(obj, device) -> CSR:
    """A table the REFERENCE saved (a pickled torch_sparse.SparseTensor: storage fields _row/_rowptr/_col/_value) or
    a torch sparse tensor -> CSR."""
    if isinstance(obj, torch.Tensor):
        from .graph import csr_from_sparse
        return csr_from_sparse(obj, device)
    st = getattr(obj, "storage", obj)
    st = st.__dict__ if hasattr(st, "__dict__") else st
    col, val = st.get("_col"), st.get("_value")
    row, rowptr = st.get("_row"), st.get("_rowptr")
    sizes = st.get("_sparse_sizes")
    if col is None or (row is None and rowptr is None):
        raise TypeError("unrecognised PPR cache file (neither an lpformer_b200 table nor a torch_sparse.SparseTensor)")
    n = int(sizes[0]) if sizes is not None else int(rowptr.numel() - 1)
    if row is None:
        row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    if val is None:
        val = torch.ones(col.numel())
    return csr_from_coo(row, col, val.float(), n, device=device)


def compute_ppr(edge_index: torch.Tensor, num_nodes: int, alpha: float, eps: float, device) -> CSR:
    """util/calc_ppr_scores.py:103-127 + :221-241: coalesce the edges (directed as given), push from every source,
    table sorted by (row, col) in fp32."""
    device = torch.device(device)
    if device.type == "cuda":
        from . import ppr as gpu_ppr
        return gpu_ppr.get_ppr_matrix(edge_index.to(device), num_nodes, alpha, eps)
    from . import synthetic
    ei = edge_index.to("cpu", torch.int64)
    key = torch.unique(ei[0] * num_nodes + ei[1])
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(torch.bincount(key // num_nodes, minlength=num_nodes), 0)
    rowptr, col, val = synthetic.ppr_push(indptr.numpy(), (key % num_nodes).to(torch.int32).numpy(), alpha, eps)
    return CSR(torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val), num_nodes)


def get_ppr(dataset: str, edge_index: torch.Tensor, num_nodes: int, alpha: float, eps: float, is_val: bool,
            device="cpu", cache_dir: Optional[str] = None, cache: bool = True) -> CSR:
    """util/calc_ppr_scores.py:244-270: load the table if its cache file exists, otherwise compute and save it."""
    path = ppr_cache_path(dataset, alpha, eps, is_val, cache_dir)
    if cache and os.path.isfile(path):
        try:
            obj = torch.load(path, map_location="cpu", weights_only=False)
        except (ModuleNotFoundError, AttributeError, pickle.UnpicklingError):
            obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)
        if isinstance(obj, dict) and obj.get("format") == "lpformer_b200.csr":
            if int(obj["n"]) != num_nodes:
                raise ValueError(f"{path}: table of {obj['n']} nodes, graph has {num_nodes}")
            return CSR(obj["rowptr"].to(device), obj["col"].to(device), obj["val"].to(device), num_nodes)
        return _csr_from_reference_file(obj, device)
    table = compute_ppr(edge_index, num_nodes, alpha, eps, device)
    if cache:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"format": "lpformer_b200.csr", "n": num_nodes, "alpha": alpha, "eps": eps, "rowptr": table.rowptr.cpu(),
                    "col": table.col.cpu(), "val": table.val.cpu()}, path)
    return table


# ------------------------------------------------------------------------------------------------------------------
# Planetoid-style fixed splits (Cora / Citeseer / Pubmed text files): util/read_datasets.py:140-254
# ------------------------------------------------------------------------------------------------------------------
def _read_pairs(path, skip_self_loops=False, node_set=None):
    out = []
    with open(path, "r") as fh:
        for line in fh:
            sub, obj = line.strip().split("\t")
            sub, obj = int(sub), int(obj)
            if node_set is not None:
                node_set.add(sub)
                node_set.add(obj)
            if skip_self_loops and sub == obj:
                continue
            out.append((sub, obj))
    return out


def _load_heart_negatives(heart_dir, data_name):
    """util/read_datasets.py:124-131,242-248: `heart/<name>/heart_{valid,test}_samples.npy`, [P, 500, 2]."""
    out = {}
    for split in ("valid", "test"):
        with open(os.path.join(heart_dir, data_name, f"heart_{split}_samples.npy"), "rb") as f:
            out[split] = torch.from_numpy(np.load(f))
    return out["valid"], out["test"]


def read_data_planetoid(args, device, data_dir: Optional[str] = None, ppr_cache_dir: Optional[str] = None,
                        ppr_cache: bool = True):
    """util/read_datasets.py:140-254.  Files: `<data_dir>/<name>/{train,valid,test}_pos.txt`, `{valid,test}_neg.txt`
    (tab-separated pairs), `gnn_feature` (torch file holding 'entity_embedding')."""
    data_dir = data_dir or DATA_DIR
    name = args.data_name
    node_set = set()
    pos = {}
    for split in ("train", "test", "valid"):         # the reference's order; self loops count as nodes, not as edges
        pos[split] = _read_pairs(os.path.join(data_dir, name, f"{split}_pos.txt"), True, node_set)
    num_nodes = len(node_set)
    neg = {split: _read_pairs(os.path.join(data_dir, name, f"{split}_neg.txt")) for split in ("test", "valid")}

    train_pos = torch.tensor(pos["train"], dtype=torch.int64).view(-1, 2)
    train_edge = train_pos.t()
    edge_index = torch.cat((train_edge, train_edge[[1, 0]]), dim=1)
    valid_pos = torch.tensor(pos["valid"], dtype=torch.int64).view(-1, 2)
    test_pos = torch.tensor(pos["test"], dtype=torch.int64).view(-1, 2)
    valid_neg = torch.tensor(neg["valid"], dtype=torch.int64).view(-1, 2)
    test_neg = torch.tensor(neg["test"], dtype=torch.int64).view(-1, 2)

    idx = torch.randperm(train_pos.size(0))[:valid_pos.size(0)]       # same draw as the reference under the same seed
    feats = torch.load(os.path.join(data_dir, name, "gnn_feature"), map_location="cpu", weights_only=False)

    key = edge_index[0] * num_nodes + edge_index[1]
    if torch.unique(key).numel() != key.numel():
        # the reference keeps a duplicated line as the VALUE 2 in adj_mask, which its node typing then misreads as a
        # common neighbour; a 0/1 table cannot hold that, so the duplicates are merged here
        warnings.warn(f"{name}: duplicate training edges merged (the reference would mistype their endpoints)")
    data = {"dataset": name, "edge_index": edge_index.to(device), "num_nodes": num_nodes,
            "train_pos": train_pos.to(device), "train_pos_val": train_pos[idx].to(device),
            "valid_pos": valid_pos.to(device), "valid_neg": valid_neg.to(device),
            "test_pos": test_pos.to(device), "test_neg": test_neg.to(device),
            "x": feats["entity_embedding"].to(device), "full_edge_index": edge_index}
    data["adj_t"], data["adj_mask"] = _csr_pair(edge_index, None, num_nodes, device)
    data["full_adj_t"], data["full_adj_mask"] = data["adj_t"], data["adj_mask"]
    data["degree"] = degree(data["edge_index"][0], num_nodes).to(device)
    data["ppr"] = get_ppr(name, data["edge_index"], num_nodes, 0.15, args.eps, False, device, ppr_cache_dir, ppr_cache)
    data["ppr_test"] = data["ppr"]
    if getattr(args, "heart", False):
        data["valid_neg"], data["test_neg"] = _load_heart_negatives(os.path.join(data_dir, "heart"), name)
    return data


# ------------------------------------------------------------------------------------------------------------------
# OGB graphs from the raw download layout
# ------------------------------------------------------------------------------------------------------------------
def _csv(path, dtype):
    import pandas as pd
    return pd.read_csv(path, compression="gzip", header=None).values.astype(dtype, copy=False)


def read_ogb_raw(name: str, root: Optional[str] = None):
    """The graph and the edge split of `ogbl-*` from `<root>/<name with _>/`, as ogb's own loader builds them
    (ogb/io/read_graph_raw.py `read_csv_graph_raw` + ogb/linkproppred/dataset_pyg.py `get_edge_split`, an
    un-vendored dependency of the reference: requirements.txt:7):
      raw/edge.csv.gz [E, 2], raw/num-node-list.csv.gz, raw/node-feat.csv.gz (optional), raw/<extra>.csv.gz
      (collab: edge_weight, edge_year; citation2: node_year), split/<type>/{train,valid,test}.pt (dicts of arrays);
      `add_inverse_edge` graphs (collab, ddi, ppa) get every edge followed by its reverse, edge attributes repeated.
    Returns (data, split_edge): a namespace with x / edge_index / edge_weight / edge_year / num_nodes (absent ones
    None) and the dict of dicts of tensors."""
    meta = OGB_META[name]
    base = os.path.join(root or DATA_DIR, name.replace("-", "_"))
    raw = os.path.join(base, "raw")
    edge = _csv(os.path.join(raw, "edge.csv.gz"), np.int64).T
    num_nodes = int(_csv(os.path.join(raw, "num-node-list.csv.gz"), np.int64).reshape(-1)[0])
    feat_path = os.path.join(raw, "node-feat.csv.gz")
    x = None
    if os.path.exists(feat_path):
        feat = _csv(feat_path, np.float32)
        x = torch.from_numpy(np.ascontiguousarray(feat))
    extras = {}
    for key in meta["edge_files"]:
        extras[key] = _csv(os.path.join(raw, key + ".csv.gz"), np.float32 if key == "edge_weight" else np.int64)
    if meta["add_inverse_edge"]:
        dup = np.repeat(edge, 2, axis=1)
        dup[0, 1::2] = edge[1]
        dup[1, 1::2] = edge[0]
        edge = dup
        extras = {k: np.repeat(v, 2, axis=0) for k, v in extras.items()}
    data = SimpleNamespace(num_nodes=num_nodes, edge_index=torch.from_numpy(np.ascontiguousarray(edge)), x=x,
                           edge_weight=None, edge_year=None)
    for k, v in extras.items():
        setattr(data, k, torch.from_numpy(np.ascontiguousarray(v)))
    split_edge = {}
    for split in ("train", "valid", "test"):
        d = torch.load(os.path.join(base, "split", meta["split"], split + ".pt"), map_location="cpu", weights_only=False)
        split_edge[split] = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    return data, split_edge


def filter_by_year(data, split_edge, year=2007):
    """util/read_datasets.py:257-279 (ogbl-collab): training edges from `year` on, the graph rebuilt from them as an
    undirected, coalesced edge list whose weights are summed."""
    tr = split_edge["train"]
    keep = (tr["year"] >= year).nonzero(as_tuple=False).reshape(-1)
    tr["edge"], tr["weight"], tr["year"] = tr["edge"][keep], tr["weight"][keep], tr["year"][keep]
    ei, w = to_undirected(tr["edge"].t(), tr["weight"], data.num_nodes)
    data.edge_index = ei
    data.edge_weight = w.unsqueeze(-1)
    return data, split_edge


def read_data_ogb(args, device, data_dir: Optional[str] = None, ppr_cache_dir: Optional[str] = None,
                  ppr_cache: bool = True, dataset=None):
    """util/read_datasets.py:20-135.  `dataset` may be a (data, split_edge) pair already in memory; otherwise the raw
    files under `<data_dir>/ogbl_<name>/` are read."""
    data_dir = data_dir or DATA_DIR
    name = args.data_name
    data, split_edge = dataset if dataset is not None else read_ogb_raw(name, data_dir)
    if "collab" in name:
        data, split_edge = filter_by_year(data, split_edge)
    n = int(data.num_nodes)
    obj = {"dataset": name, "num_nodes": n}
    edge_index = data.edge_index.to(device)

    if name != "ogbl-citation2":
        obj["train_pos"] = split_edge["train"]["edge"].to(device)
        obj["valid_pos"] = split_edge["valid"]["edge"].to(device)
        obj["valid_neg"] = split_edge["valid"]["edge_neg"].to(device)
        obj["test_pos"] = split_edge["test"]["edge"].to(device)
        obj["test_neg"] = split_edge["test"]["edge_neg"].to(device)
    else:
        for split, key in (("train", "train_pos"), ("valid", "valid_pos"), ("test", "test_pos")):
            s = split_edge[split]
            obj[key] = torch.stack([s["source_node"], s["target_node"]], dim=-1).to(device)
        obj["valid_neg"] = split_edge["valid"]["target_node_neg"].to(device)
        obj["test_neg"] = split_edge["test"]["target_node_neg"].to(device)

    heart_dir = os.path.join(data_dir, "heart")
    heart = bool(getattr(args, "heart", False))
    if heart and "ppa" in name:          # HeaRT scores a fixed subsample of ogbl-ppa's positives
        val_ix = torch.load(os.path.join(heart_dir, name, "valid_samples_index.pt"), map_location="cpu", weights_only=False)
        test_ix = torch.load(os.path.join(heart_dir, name, "test_samples_index.pt"), map_location="cpu", weights_only=False)
        obj["valid_pos"] = obj["valid_pos"][torch.as_tensor(val_ix).to(device), :]
        obj["test_pos"] = obj["test_pos"][torch.as_tensor(test_ix).to(device), :]

    idx = torch.randperm(obj["train_pos"].size(0))[:obj["valid_pos"].size(0)]
    obj["train_pos_val"] = obj["train_pos"][idx.to(device)]

    if getattr(data, "x", None) is not None:
        obj["x"] = data.x.to(device).to(torch.float)
    else:                                # ogbl-ddi: a free embedding table kept in the dict
        obj["x"] = torch.nn.Parameter(torch.zeros(n, args.dim).to(device))
        torch.nn.init.xavier_uniform_(obj["x"])

    ew = getattr(data, "edge_weight", None)
    edge_weight = ew.to(device).to(torch.float).view(-1) if ew is not None else torch.ones(edge_index.size(1), device=device)

    adj_t, mask = _csr_pair(edge_index, edge_weight, n, device)
    if name == "ogbl-citation2":         # directed: symmetrised for the GCN and for the node typing (:88-93)
        sym = torch.cat([edge_index, edge_index.flip(0)], 1)
        adj_t, mask = _csr_pair(sym, torch.cat([edge_weight, edge_weight]), n, device)
        # SparseTensor.to_symmetric(reduce='sum') ADDS the two directions of an edge present both ways and coalesce()
        # merges duplicates the same way, which is what summing duplicate COO entries gives
    else:                                # undirected already; adj_mask = adj_t.to_symmetric() pattern
        sym = torch.cat([edge_index, edge_index.flip(0)], 1)
        _, mask = _csr_pair(sym, None, n, device)
    obj["adj_t"], obj["adj_mask"] = adj_t, mask

    if getattr(args, "use_val_in_test", False):
        val_ei = to_undirected(split_edge["valid"]["edge"].t(), None, n).to(device)
        full_ei = torch.cat([edge_index, val_ei], dim=-1)
        obj["full_edge_index"] = full_ei
        full_w = torch.cat([edge_weight, torch.ones(val_ei.size(1), device=device)])
        obj["full_adj_t"], obj["full_adj_mask"] = _csr_pair(full_ei, full_w, n, device)
    else:
        obj["full_adj_t"], obj["full_adj_mask"], obj["full_edge_index"] = obj["adj_t"], obj["adj_mask"], edge_index

    obj["degree"] = degree(edge_index[0], n).to(device)
    if getattr(args, "use_val_in_test", False):
        obj["degree_test"] = degree(obj["full_edge_index"][0], n).to(device)

    obj["ppr"] = get_ppr(name, edge_index, n, 0.15, args.eps, False, device, ppr_cache_dir, ppr_cache)
    if getattr(args, "use_val_in_test", False):
        obj["ppr_test"] = get_ppr(name, obj["full_edge_index"], n, 0.15, args.eps, True, device, ppr_cache_dir, ppr_cache)
    else:
        obj["ppr_test"] = obj["ppr"]

    if heart:
        v, t = _load_heart_negatives(heart_dir, name)
        obj["valid_neg"], obj["test_neg"] = v.to(device), t.to(device)
        if "ddi" in name:                # a quarter of the validation queries (:133-138)
            num = obj["valid_pos"].size(0) // 4
            idx = torch.randperm(obj["valid_pos"].size(0))[:num].to(device)
            obj["valid_pos"], obj["valid_neg"] = obj["valid_pos"][idx], obj["valid_neg"][idx]
            obj["train_pos_val"] = obj["train_pos_val"][idx]
    return obj


def read_data(args, device, **kw):
    """run.py:154-157: OGB names go to `read_data_ogb`, everything else is a Planetoid-style directory."""
    return read_data_ogb(args, device, **kw) if "ogbl" in args.data_name else read_data_planetoid(args, device, **kw)
