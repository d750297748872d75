"""PPR precompute on the GPU — the reference's offline tool (util/calc_ppr_scores.py) as a device-side step.

`get_ppr_matrix(edge_index, num_nodes, alpha, eps)` mirrors the reference function of the same name
(util/calc_ppr_scores.py:103-127: coalesce the edges, CSR, Andersen push for every source) but returns the table
the model consumes — the sorted (row, col) fp32 matrix the reference builds afterwards (:221-241) and loads as
`data['ppr']` (util/read_datasets.py:122-129) — as a CSR on the device.  Values are bit-identical to the
reference's numba kernel (tests compare with the host port, which is itself pinned to the numba kernel).

The push runs in lpf_ppr_push (csrc/ppr_push_gpu.cu: one warp per source, hash table per warp); this module owns the
plumbing around it: scratch sizing, the output pool (grown and re-run when it overflows), and the one device
sort by (row, col) that turns the emitted entries into the CSR.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream
from .graph import CSR


def csr_from_edge_index(edge_index: torch.Tensor, num_nodes: int):
    """Coalesced, sorted CSR (indptr int64, indices int32) of a [2, E] edge list on its device."""
    key = torch.unique(edge_index[0].to(torch.int64) * num_nodes + edge_index[1].to(torch.int64))
    row, col = key // num_nodes, (key % num_nodes).to(torch.int32)
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=edge_index.device)
    indptr[1:] = torch.cumsum(torch.bincount(row, minlength=num_nodes), 0)
    return indptr, col.contiguous()


def ppr_push(indptr: torch.Tensor, indices: torch.Tensor, alpha: float = 0.15, eps: float = 5e-5,
             nwarps: int | None = None, cap: int | None = None, scratch_limit_bytes: int = 8 << 30,
             first_pass_slots: int = 1 << 12) -> CSR:
    """PPR table of a CSR graph (device tensors: indptr int64 [n+1], indices int32, sorted columns, no self loops)
    as a sorted CSR (rowptr int64, col int32, val fp32) on the same device."""
    if not indptr.is_cuda or not indices.is_cuda:
        raise _lib.LpfError("ppr_push needs CUDA tensors: lpformer_b200 has no CPU fallback (the host tool is "
                            "lpformer_b200.synthetic.ppr_push)")
    lib = _lib.load()
    dev = indptr.device
    n = indptr.numel() - 1
    indptr = indptr.to(torch.int64).contiguous()
    indices = indices.to(torch.int32).contiguous()
    # worst-case keys per source: 1 + 1 / (alpha eps) (lpf_ppr_push_slots), and never more than the n nodes there are
    slots_n = 64
    while slots_n < 2 * (n + 1):
        slots_n <<= 1
    slots_max = lib.lpf_ppr_push_slots(float(alpha), float(eps))
    slots_max = slots_n if slots_max < 0 else min(slots_max, slots_n)
    if slots_max > (1 << 25):
        raise _lib.LpfError(f"eps = {eps} on {n} nodes needs more than 2^25 hash slots per source: use the host tool "
                            "(lpformer_b200.synthetic.ppr_push) for this table")

    def warps_for(slots):
        w = 148 * 16 if nwarps is None else int(nwarps)
        while w > 4 and lib.lpf_ppr_push_scratch_bytes(slots, w) > scratch_limit_bytes:
            w //= 2
        return max(4, (w + 3) // 4 * 4)

    # slots_max covers the worst case (1 + 1/(alpha eps) touched nodes); almost every source touches far fewer, so all
    # sources run with small tables first (L2-friendly) and only those whose table filled up run again with the full size
    slots1 = min(slots_max, int(first_pass_slots))
    cap = int(cap if cap is not None else max(1 << 16, 8 * n))
    ovf = torch.empty(max(1, n), dtype=torch.int32, device=dev)
    while True:
        row = torch.empty(cap, dtype=torch.int32, device=dev)
        col = torch.empty(cap, dtype=torch.int32, device=dev)
        val = torch.empty(cap, dtype=torch.float32, device=dev)
        cursor = torch.zeros(2, dtype=torch.int64, device=dev)
        status = torch.zeros(2, dtype=torch.int32, device=dev)
        w1 = warps_for(slots1)
        scratch = torch.empty(lib.lpf_ppr_push_scratch_bytes(slots1, w1), dtype=torch.uint8, device=dev)
        call("lpf_ppr_push", ptr(indptr), ptr(indices), n, float(alpha), float(eps), 0, n, None, ptr(ovf), slots1, w1,
             ptr(scratch), ptr(row), ptr(col), ptr(val), cap, ptr(cursor), ptr(status), stream())
        n_ovf = int(status[1])
        if n_ovf and slots1 < slots_max:
            w2 = warps_for(slots_max)
            scratch = torch.empty(lib.lpf_ppr_push_scratch_bytes(slots_max, w2), dtype=torch.uint8, device=dev)
            cursor[1] = 0
            status[1] = 0
            call("lpf_ppr_push", ptr(indptr), ptr(indices), n, float(alpha), float(eps), 0, n_ovf, ptr(ovf), None,
                 slots_max, w2, ptr(scratch), ptr(row), ptr(col), ptr(val), cap, ptr(cursor), ptr(status), stream())
            n_ovf = int(status[1])
        if n_ovf:
            raise _lib.LpfError("lpf_ppr_push: a per-source table filled up (is the graph simple, without self loops?)")
        nnz = int(cursor[0])
        if not int(status[0]) and nnz <= cap:
            break
        cap = max(2 * cap, nnz + 1024)          # the cursor kept counting: the exact size is known now
    del scratch
    row, col, val = row[:nnz], col[:nnz], val[:nnz]
    order = torch.argsort(row.to(torch.int64) * n + col.to(torch.int64))
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(row.to(torch.int64), minlength=n), 0)
    return CSR(rowptr, col[order].contiguous(), val[order].contiguous(), n)


def get_ppr_matrix(edge_index: torch.Tensor, num_nodes: int, alpha: float = 0.15, eps: float = 5e-5) -> CSR:
    """Reference util/calc_ppr_scores.py:103-127 (same arguments); returns the sorted PPR table as a device CSR."""
    indptr, indices = csr_from_edge_index(edge_index, num_nodes)
    return ppr_push(indptr, indices, alpha, eps)


def to_sparse_coo(ppr: CSR) -> torch.Tensor:
    """The layout the reference keeps in data['ppr'] (util/read_datasets.py:122-129): coalesced fp32 sparse COO."""
    n = ppr.n
    counts = ppr.rowptr[1:] - ppr.rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(n, device=ppr.col.device), counts)
    return torch.sparse_coo_tensor(torch.stack([row, ppr.col.to(torch.int64)]), ppr.val, (n, n), is_coalesced=True)
