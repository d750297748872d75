"""Seeded synthetic inputs of the shapes BASELINE.json names (no datasets / network here).

Degree-skewed undirected simple graphs (Chung-Lu with a shifted power law), PPR tables by the
reference's push algorithm (host tool in csrc/ppr_push.cpp), node features ~ N(0,1), and
candidate-link workloads (citation2-style: 1 held-out positive + K uniform-random negatives
sharing the query's source; HeaRT-style: K negatives per positive, half of them corrupted
towards 2-hop neighbours).  Everything is generated on the host with numpy.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .graph import CSR

# name -> shape + hyper-parameters of reference scripts/replicate_heart.sh
CONFIGS = {
    "cora": dict(n=2708, e=5278, feat=1433, dim=256, gnn_layers=1, residual=False, layer_norm=False, relu=False,
                 thresh_cn=0, thresh_1hop=1e-2, thresh_non1hop=1e-2, eps=1e-7, batch=16384, negs=500),
    "collab": dict(n=235868, e=1285465, feat=128, dim=128, gnn_layers=3, residual=False, layer_norm=True, relu=True,
                   thresh_cn=0, thresh_1hop=1e-4, thresh_non1hop=1e-2, eps=5e-5, batch=32768, negs=500),
    "ddi": dict(n=4267, e=1334889, feat=0, dim=256, gnn_layers=3, residual=False, layer_norm=True, relu=True,
                thresh_cn=0, thresh_1hop=1e-2, thresh_non1hop=1, eps=5e-6, batch=8192, negs=500),
    "ppa": dict(n=576289, e=30326273, feat=58, dim=64, gnn_layers=3, residual=True, layer_norm=True, relu=True,
                thresh_cn=0, thresh_1hop=1e-4, thresh_non1hop=1e-2, eps=5e-5, batch=32768, negs=500),
    "citation2": dict(n=2927963, e=30561187, feat=128, dim=64, gnn_layers=3, residual=True, layer_norm=True,
                      relu=True, thresh_cn=0, thresh_1hop=1e-3, thresh_non1hop=1e-2, eps=2.5e-3, batch=32768,
                      negs=1000),
}


def train_args_of(cfg):
    keys = ("dim", "gnn_layers", "residual", "layer_norm", "relu", "thresh_cn", "thresh_1hop", "thresh_non1hop")
    return dict({k: cfg[k] for k in keys}, num_heads=1, trans_layers=1)


def _heavy_device():
    """Device for the sort / unique / searchsorted steps of graph generation: the GPU when there is
    one (seconds instead of minutes at 30M edges); results are identical on either device because
    the random numbers always come from numpy on the host."""
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def chung_lu_edges(n, m, seed, gamma=2.5, i0=None):
    """[2, m'] undirected simple edges (lo < hi), m' <= m (a little lower after dedup on dense shapes).
    Expected degree of node i ~ (i + i0)^(-1/(gamma-1)); i0 caps the maximum degree."""
    rng = np.random.default_rng(seed)
    dev = _heavy_device()
    expo = 1.0 / (gamma - 1.0)
    i0 = max(1.0, n / 30000.0) if i0 is None else i0
    w = (np.arange(n, dtype=np.float64) + i0) ** -expo
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    cdf_t = torch.from_numpy(cdf).to(dev)
    keys = torch.zeros(0, dtype=torch.int64, device=dev)
    need = m
    dense = m > 0.05 * n * (n - 1) / 2
    for _ in range(40):
        k = int(need * (2.5 if dense else 1.15)) + 1024
        a = torch.searchsorted(cdf_t, torch.from_numpy(rng.random(k)).to(dev)).clamp_(max=n - 1)
        b = torch.searchsorted(cdf_t, torch.from_numpy(rng.random(k)).to(dev)).clamp_(max=n - 1)
        keep = a != b
        lo, hi = torch.minimum(a, b)[keep], torch.maximum(a, b)[keep]
        keys = torch.unique(torch.cat([keys, lo * n + hi]))
        if keys.numel() >= m:
            break
        need = m - keys.numel()
    if keys.numel() > m:
        sel = np.sort(rng.permutation(keys.numel())[:m])
        keys = keys[torch.from_numpy(sel).to(dev)]
    # relabel nodes randomly so that node id carries no degree information
    perm = torch.from_numpy(rng.permutation(n).astype(np.int64)).to(dev)
    out = torch.stack([perm[keys // n], perm[keys % n]])
    return out.cpu().numpy()


def symmetric_csr(edges, n):
    """Host numpy CSR (indptr int64, indices int32 ascending) of the symmetrised edge list."""
    dev = _heavy_device()
    e = torch.from_numpy(np.ascontiguousarray(edges)).to(dev)
    key = torch.unique(torch.cat([e[0] * n + e[1], e[1] * n + e[0]]))
    row = key // n
    col = (key - row * n).to(torch.int32)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return indptr.cpu().numpy(), col.cpu().numpy()


def ppr_push(indptr, indices, alpha, eps, nthreads=0):
    """PPR CSR (rowptr int64, col int32, val fp32) by the host push tool (reference algorithm)."""
    lib = _lib.load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    n = indptr.size - 1
    nnz = C.c_int64(0)
    h = lib.lpf_ppr_push_host(indptr.ctypes.data, indices.ctypes.data, n, float(alpha), float(eps), int(nthreads),
                              C.addressof(nnz))
    if not h:
        raise _lib.LpfError("lpf_ppr_push_host failed")
    rowptr = np.empty(n + 1, dtype=np.int64)
    col = np.empty(nnz.value, dtype=np.int32)
    val = np.empty(nnz.value, dtype=np.float32)
    rc = lib.lpf_ppr_push_host_fetch(h, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data)
    if rc != 0:
        raise _lib.LpfError("lpf_ppr_push_host_fetch failed")
    return rowptr, col, val


@dataclass
class SyntheticGraph:
    name: str
    cfg: dict
    n: int
    edges: np.ndarray          # [2, E] lo<hi, in the graph
    heldout: np.ndarray        # [2, Q] positives NOT in the graph
    indptr: np.ndarray
    indices: np.ndarray
    ppr: tuple                 # (rowptr, col, val) numpy
    x: np.ndarray              # [N, F] fp32

    def data_dict(self, device):
        """The reference's `data` dict with the sparse tables already in CSR form (LinkTransformer
        accepts CSR objects wherever the reference passes sparse tensors)."""
        t = torch.from_numpy
        adj = CSR(t(self.indptr).to(device), t(self.indices).to(device), None, self.n)
        adj_w = CSR(adj.rowptr, adj.col, torch.ones(adj.col.numel(), dtype=torch.float32, device=device), self.n)
        ppr = CSR(t(self.ppr[0]).to(device), t(self.ppr[1]).to(device), t(self.ppr[2]).to(device), self.n)
        x = t(self.x).to(device)
        return {"x": x, "adj_t": adj_w, "adj_mask": adj, "ppr": ppr, "full_adj_t": adj_w, "full_adj_mask": adj,
                "ppr_test": ppr}

    def stats(self):
        deg = np.diff(self.indptr)
        npp = np.diff(self.ppr[0])
        return {"nodes": int(self.n), "edges": int(self.edges.shape[1]), "deg_mean": float(deg.mean()),
                "deg_median": float(np.median(deg)), "deg_max": int(deg.max()), "ppr_nnz_per_row": float(npp.mean()),
                "ppr_row_max": int(npp.max())}


def make_graph(name, seed=0, scale=1.0, heldout=4096, nthreads=0) -> SyntheticGraph:
    cfg = dict(CONFIGS[name])
    n = max(64, int(round(cfg["n"] * scale)))
    e = max(64, int(round(cfg["e"] * scale)))
    e = min(e, n * (n - 1) // 2 - heldout)
    all_edges = chung_lu_edges(n, e + heldout, seed)
    rng = np.random.default_rng(seed + 1)
    pick = rng.permutation(all_edges.shape[1])
    held = all_edges[:, pick[:heldout]]
    edges = all_edges[:, np.sort(pick[heldout:])]
    indptr, indices = symmetric_csr(edges, n)
    ppr = ppr_push(indptr, indices, 0.15, cfg["eps"], nthreads)
    feat = cfg["feat"] if cfg["feat"] > 0 else cfg["dim"]     # ddi: random [N, dim] (read_datasets.py:76-77)
    x = np.random.default_rng(seed + 2).standard_normal((n, feat), dtype=np.float32)
    return SyntheticGraph(name, cfg, n, edges, held, indptr, indices, ppr, x)


def citation2_queries(g: SyntheticGraph, num_queries, negs, seed=1):
    """[2, Q*(1+negs)] int64: per query the held-out positive (a,b) then `negs` links (a, random node)
    (reference train/testing.py:20-23: source.repeat(1000) x negative targets)."""
    rng = np.random.default_rng(seed)
    q = g.heldout[:, rng.integers(0, g.heldout.shape[1], num_queries)]
    flip = rng.random(num_queries) < 0.5
    src = np.where(flip, q[1], q[0])
    pos = np.where(flip, q[0], q[1])
    tgt = np.concatenate([pos[:, None], rng.integers(0, g.n, (num_queries, negs))], axis=1)
    return np.stack([np.repeat(src, 1 + negs), tgt.reshape(-1)]).astype(np.int64)


def heart_queries(g: SyntheticGraph, num_pos, negs, seed=1):
    """HeaRT-style: per positive (a,b), `negs` negatives; half corrupt the target with a uniform node,
    half with a 2-hop neighbour of the source (hard negatives)."""
    rng = np.random.default_rng(seed)
    q = g.heldout[:, rng.integers(0, g.heldout.shape[1], num_pos)]
    src = np.repeat(q[0], 1 + negs).reshape(num_pos, 1 + negs)
    tgt = np.empty((num_pos, 1 + negs), dtype=np.int64)
    tgt[:, 0] = q[1]
    half = negs // 2
    tgt[:, 1:1 + half] = rng.integers(0, g.n, (num_pos, half))
    deg = np.diff(g.indptr)
    for i in range(num_pos):   # 2-hop walk from the source
        a = q[0, i]
        k = negs - half
        if deg[a] == 0:
            tgt[i, 1 + half:] = rng.integers(0, g.n, k)
            continue
        n1 = g.indices[g.indptr[a] + rng.integers(0, deg[a], k)]
        d1 = np.maximum(deg[n1], 1)
        n2 = g.indices[np.minimum(g.indptr[n1] + rng.integers(0, 1 << 30, k) % d1, g.indices.size - 1)]
        tgt[i, 1 + half:] = n2
    return np.stack([src.reshape(-1), tgt.reshape(-1)]).astype(np.int64)
