"""Tensor-level wrappers over the C-ABI (one function per entry point).

All tensors must live on the CUDA device; outputs are allocated with torch (device memory
plumbing) and every launch goes to torch's current stream.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import EPI_NONE, EPI_RELU, EPI_SIGMOID, MODE, call, ptr, require_cuda, stream
from .graph import CSR

TYPE_NAMES = ("cn", "1hop", "non1hop")


def _rowmajor(x: torch.Tensor) -> torch.Tensor:
    """fp32 2-D tensor with unit stride in the last dimension (row stride arbitrary)."""
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    return x


def gemm(A, W, bias=None, bias_scale=1.0, out=None, epilogue=EPI_NONE):
    """out[M,N] = epi(A[M,K] @ W[N,K]^T + bias_scale * bias)."""
    require_cuda(A, W, bias, out)
    A, W = _rowmajor(A), _rowmajor(W)
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K:
        raise ValueError(f"gemm: A is [{M},{K}] but W is {list(W.shape)}")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32:
        raise ValueError("gemm: bad `out`")
    if bias is not None:
        bias = bias.float().contiguous()
    call("lpf_gemm", ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), float(bias_scale), ptr(out),
         out.stride(0) if M > 0 else N, M, N, K, epilogue, stream(), meta=(M, N, K))
    return out


_PACKED = {}          # (data_ptr, version, shape, stride) -> (packed image, weight kept alive)
USE_PACKED_ROWS = os.environ.get("LPF_PACKED_ROWS", "1") != "0"   # one-pass selection over the packed link rows
GEMM_BACKEND = os.environ.get("LPF_GEMM_BACKEND", "tc")   # "tc": tcgen05 3xTF32 kernel (lpf_gemm_tc);  "simt": fp32 FFMA kernel (lpf_gemm)


def pack_weight(W):
    """Pre-split (hi/lo tf32) pre-swizzled image of an nn.Linear weight [N,K] for lpf_gemm_tc; cached per
    (storage, version) so it is built once per weight."""
    key = (W.data_ptr(), W._version, tuple(W.shape), W.stride(0), W.device.index)
    hit = _PACKED.get(key)
    if hit is not None:
        return hit[0]
    require_cuda(W)
    Wc = _rowmajor(W.detach())
    N, K = Wc.shape
    nbytes = _lib.load().lpf_pack_weight_bytes(N, K)
    packed = torch.empty(nbytes // 4, dtype=torch.float32, device=W.device)
    call("lpf_pack_weight", ptr(Wc), Wc.stride(0), N, K, ptr(packed), stream())
    if len(_PACKED) > 512:
        _PACKED.clear()
    _PACKED[key] = (packed, W)
    return packed


HEADS_F16 = os.environ.get("LPF_HEADS", "f16") != "tf32"   # d = 64 heads: fp16-split operands (lpf_link_heads_f16) or 3xTF32


def pow2_scale(bound, target_exp=15):
    """Largest power of two s with bound * s < 2^target_exp (1.0 for a zero bound)."""
    import math
    if not (bound > 0.0) or not math.isfinite(bound):
        return 1.0
    return 2.0 ** (target_exp - 1 - math.floor(math.log2(bound)))


def pack_weight_f16(W):
    """(image, scale): fp16 hi / lo images of an nn.Linear weight [N, K] in the swizzled K-major layout of
    tcgen05.mma.kind::f16 (lpf_pack_weight_f16), scaled by the power of two that puts max|W| into [2^14, 2^15)."""
    require_cuda(W)
    Wc = _rowmajor(W.detach())
    N, K = Wc.shape
    scale = pow2_scale(float(Wc.abs().max()))
    nbytes = _lib.load().lpf_pack_weight_f16_bytes(N, K)
    packed = torch.empty(nbytes // 4, dtype=torch.float32, device=W.device)
    call("lpf_pack_weight_f16", ptr(Wc), Wc.stride(0), N, K, float(scale), ptr(packed), stream())
    return packed, scale


def heads_call(links, bs, idx, n, X, c, zb, ld_zb, prob, logits, n_dev, sched, st):
    """One launch of the fused heads with the operand set `c` (model._head_consts): the fp16-split kernel when the
    consts carry its images (d = 64), the 3xTF32 kernel otherwise.  idx / zb / n_dev / sched: raw pointers or None."""
    d = X.shape[1]
    bf = X.dtype == torch.bfloat16
    if "w1h" in c and X.data_ptr() % 16 == 0 and (X.stride(0) * X.element_size()) % 16 == 0:
        call("lpf_link_heads_f16", ptr(links), bs, idx, n, ptr(X), int(bf), X.stride(0), X.shape[0], d, ptr(c["w1h"]), c["inv_sw1"], ptr(c["b1"]),
             ptr(c["ln_w_s"]), ptr(c["ln_b_s"]), ptr(c["w23h"]), c["inv_s3"], ptr(c["c3"]) if zb is None else None, zb, ld_zb,
             ptr(c["ws2"]), ptr(c["bs2"]), ptr(prob), int(logits), n_dev, st, meta=(n,),
             label=None if zb is None else "non-empty links (per-row offsets)")
    else:
        if bf:
            raise _lib.LpfError("bf16 node tables need the fp16-split heads (d = 64, 16-byte aligned rows)")
        call("lpf_link_heads_tc", ptr(links), bs, idx, n, ptr(X), X.stride(0), d, ptr(c["w1p"]), ptr(c["b1"]),
             ptr(c["ln_w"]), ptr(c["ln_b"]), ptr(c["w23p"]), ptr(c["c3"]) if zb is None else None, zb, ld_zb,
             ptr(c["ws2"]), ptr(c["bs2"]), ptr(prob), int(logits), n_dev, sched, st, meta=(n,),
             label=None if zb is None else "non-empty links (per-row offsets)")


def linear(A, W, bias=None, bias_scale=1.0, out=None, epilogue=EPI_NONE):
    """out[M,N] = epi(A @ W^T + bias_scale*bias) on the tensor cores (rows of W in chunks of <= 256)."""
    if GEMM_BACKEND != "tc":
        return gemm(A, W, bias, bias_scale, out, epilogue)
    require_cuda(A, W, bias, out)
    A = _rowmajor(A)
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K:
        raise ValueError(f"linear: A is [{M},{K}] but W is {list(W.shape)}")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32:
        raise ValueError("linear: bad `out`")
    if bias is not None:
        bias = bias.detach().float().contiguous()
    for n0 in range(0, N, 256):
        n1 = min(N, n0 + 256)
        Wc = W if (n0 == 0 and n1 == N) else W[n0:n1]
        o = out if (n0 == 0 and n1 == N) else out[:, n0:n1]
        b = None if bias is None else (bias if (n0 == 0 and n1 == N) else bias[n0:n1])
        call("lpf_gemm_tc", ptr(A), A.stride(0), ptr(pack_weight(Wc)), ptr(b), float(bias_scale), ptr(o),
             o.stride(0) if M > 0 else N, M, n1 - n0, K, epilogue, None, stream(), meta=(M, n1 - n0, K))
    return out


def layernorm_act(x, weight, bias, relu=False, residual=None, out=None, n=None):
    """out = residual + act(LayerNorm(x[:, :n])) row-wise; weight=None skips the norm."""
    require_cuda(x, weight, bias, residual, out)
    x = _rowmajor(x)
    rows = x.shape[0]
    n = x.shape[1] if n is None else n
    if out is None:
        out = torch.empty((rows, n), dtype=torch.float32, device=x.device)
    if residual is not None:
        residual = _rowmajor(residual)
    call("lpf_layernorm_act", ptr(x), x.stride(0), ptr(weight), ptr(bias), ptr(residual),
         residual.stride(0) if residual is not None else 0, ptr(out), out.stride(0), rows, n, int(relu), None, stream())
    return out


def gather_links(links, X, want_sum=True, want_prod=True, out_sum=None, out_prod=None, idx=None):
    """xsum = X[a]+X[b], xprod = X[a]*X[b] for every link, or for the batch positions listed in idx."""
    require_cuda(links, X, idx)
    X = _rowmajor(X)
    bs, d = links.shape[1], X.shape[1]
    n = bs if idx is None else idx.numel()
    if want_sum and out_sum is None:
        out_sum = torch.empty((n, d), dtype=torch.float32, device=X.device)
    if want_prod and out_prod is None:
        out_prod = torch.empty((n, d), dtype=torch.float32, device=X.device)
    call("lpf_gather_links", ptr(links), bs, ptr(idx), n, ptr(X), X.stride(0), d, ptr(out_sum),
         out_sum.stride(0) if out_sum is not None else 0, ptr(out_prod),
         out_prod.stride(0) if out_prod is not None else 0, None, int(X.dtype == torch.bfloat16), stream())
    return out_sum, out_prod


def scatter_rows(src, idx, dst, fill_row=None):
    """dst[:, :] = fill_row (if given), then dst[idx[j], :] = src[j, :]."""
    require_cuda(src, idx, dst, fill_row)
    n = 0 if idx is None or src is None else idx.numel()
    d = dst.shape[1]
    call("lpf_scatter_rows", ptr(src), src.stride(0) if n > 0 else d, ptr(idx), n, ptr(dst), dst.stride(0), dst.shape[0], d,
         ptr(fill_row), stream())
    return dst


@dataclass
class Selection:
    """Selected (link, node) pairs of one batch, type-major (CN | 1-hop | >1-hop), each type
    sorted by (link, node): the concatenation order of reference link_transformer.py:161."""
    mode: str
    bs: int
    ptr: torch.Tensor            # int64 [3*bs+1]
    node: torch.Tensor           # int32 [S]
    src_ppr: torch.Tensor        # fp32 [S]
    tgt_ppr: torch.Tensor        # fp32 [S]
    link: Optional[torch.Tensor]  # int32 [S] or None
    bounds: tuple                # (0, S_cn, S_cn+S_1hop, S) python ints
    nz: Optional[torch.Tensor] = None   # int32 [M']: batch positions of the links with a non-empty set (any order)

    @property
    def total(self) -> int:
        return self.bounds[3]

    def type_range(self, t: int):
        return self.bounds[t], self.bounds[t + 1]

    def counts(self) -> torch.Tensor:
        """int64 [3, bs] set sizes."""
        p = self.ptr
        return (p[1:] - p[:-1]).view(3, self.bs)


def merge_selections(near: Selection, far: Selection) -> Selection:
    """CN and 1-hop sets of `near` with the >1-hop sets of `far` (two selections of the same batch over different
    adjacency tables: a caller-supplied one decides CN / 1-hop, the stored one the >1-hop set — reference
    models/link_transformer.py:226-254 vs :443-447).  Slow path (torch ops, one host sync)."""
    bs, dev = near.bs, near.node.device
    cn, hop = near.counts()[0], near.counts()[1]
    far_cnt = far.counts()[2]
    counts = torch.cat((cn, hop, far_cnt))
    p = torch.zeros(3 * bs + 1, dtype=torch.int64, device=dev)
    p[1:] = torch.cumsum(counts, 0)
    n0, n1 = near.bounds[1], near.bounds[2]
    f0, f1 = far.bounds[2], far.bounds[3]
    cat = lambda a, b: torch.cat((a[:n1], b[f0:f1]))    # noqa: E731
    link = None if near.link is None or far.link is None else cat(near.link, far.link)
    nz = torch.nonzero((cn + hop + far_cnt) > 0).reshape(-1).to(torch.int32)
    return Selection("all", bs, p, cat(near.node, far.node), cat(near.src_ppr, far.src_ppr), cat(near.tgt_ppr, far.tgt_ppr),
                     link, (0, n0, n1, n1 + (f1 - f0)), nz)


def drop_packed_weights():
    """Forgets every cached tensor-core weight image (LinkTransformer.invalidate_weights)."""
    _PACKED.clear()


def links_tensor(batch, device) -> torch.Tensor:
    """int64 [2,BS] contiguous on `device` (reference forward() moves the batch, :98)."""
    b = torch.as_tensor(batch)
    if b.dim() != 2 or b.shape[0] != 2:
        raise ValueError("batch must be a [2, BS] tensor of (source, target) node ids")
    return b.to(device=device, dtype=torch.int64).contiguous()


def pick_select_algo(adj: CSR, ppr: CSR, th_1hop, th_non1hop, mode: str) -> int:
    """Intersection-driven kernel when its preconditions hold (thresholds gating PPR-derived sets > 0 and
    all PPR values in (0,1], checked once per table), group width by mean row length; else the generic one."""
    ok = mode == "cn" or (th_1hop > 0 and (mode != "all" or th_non1hop > 0))
    if not ok:
        return _lib.ALGO_GENERIC
    if ppr.unit_range is None:
        ppr.unit_range = bool(ppr.val.numel() == 0 or (float(ppr.val.min()) > 0.0 and float(ppr.val.max()) <= 1.0))
    if not ppr.unit_range:
        return _lib.ALGO_GENERIC
    # rows walked per link: adjacency rows, and the PPR rows when the mode has PPR-derived sets — a group defers a link
    # whose shorter row exceeds 16 elements per lane to the CTA-wide kernel, so long PPR rows (ogbl-collab shape: ~325
    # entries) on a sparse graph want full warps too (measured there: every link deferred with groups of 8, 1.5 ms of
    # CTA-wide walks per 32,565-link batch)
    mean_row = adj.nnz / max(1, adj.n)
    if mode != "cn":
        mean_row = max(mean_row, ppr.nnz / max(1, ppr.n))
    return _lib.ALGO_INTERSECT32 if mean_row >= 96 else _lib.ALGO_INTERSECT8


def select(links, adj: CSR, ppr: CSR, th_cn, th_1hop, th_non1hop, mode: str, want_link=False, algo=None) -> Selection:
    """K1: count -> scan -> fill.  One host sync (reading the three totals) sizes the outputs."""
    require_cuda(links, adj.rowptr, ppr.rowptr)
    dev = links.device
    bs = links.shape[1]
    m = MODE[mode]
    st = stream()
    counts = torch.empty(3 * bs, dtype=torch.int32, device=dev)
    args = (ptr(links), bs, ptr(adj.rowptr), ptr(adj.col), ptr(ppr.rowptr), ptr(ppr.col), ptr(ppr.val),
            float(th_cn), float(th_1hop), float(th_non1hop), m,
            pick_select_algo(adj, ppr, th_1hop, th_non1hop, mode) if algo is None else algo)
    ws = torch.empty(_lib.load().lpf_select_workspace_bytes(bs) // 4, dtype=torch.int32, device=dev)
    call("lpf_select_count", *args, ptr(counts), ptr(ws), st, meta=(bs,))
    p = torch.empty(3 * bs + 1, dtype=torch.int64, device=dev)
    scratch = torch.empty(max(1, _lib.load().lpf_scan_scratch_bytes(3 * bs) // 8), dtype=torch.int64, device=dev)
    call("lpf_scan_counts", ptr(counts), 3 * bs, ptr(p), ptr(scratch), st)
    nz = torch.empty(bs, dtype=torch.int32, device=dev)
    header = torch.empty(4, dtype=torch.int64, device=dev)
    call("lpf_select_compact", ptr(p), bs, ptr(nz), ptr(header), st)
    b = header.tolist()          # the one host sync of the batch: (S_cn, S_cn+S_1hop, S, #non-empty links)
    S = b[2]
    node = torch.empty(S, dtype=torch.int32, device=dev)
    pa = torch.empty(S, dtype=torch.float32, device=dev)
    pb = torch.empty(S, dtype=torch.float32, device=dev)
    link = torch.empty(S, dtype=torch.int32, device=dev) if want_link else None
    if S > 0:
        call("lpf_select_fill", *args, ptr(p), ptr(node), ptr(pa), ptr(pb), ptr(link), ptr(ws), st, meta=(bs, S))
    return Selection(mode, bs, p, node, pa, pb, link, (0, b[0], b[1], b[2]), nz[:b[3]])


class LinkRows:
    """Packed link rows of one (adjacency, PPR) table pair (lpf_pack_link_rows): `slab` = one 128-byte line per node
    (header + the first seven 16-byte chunks of its row) and `overflow` = the rest of the longer rows."""

    def __init__(self, adj: CSR, ppr: CSR):
        require_cuda(adj.rowptr, ppr.rowptr)
        lib = _lib.load()
        dev = adj.rowptr.device
        n = adj.n
        nbytes = lib.lpf_link_rows_bytes(n, adj.nnz, ppr.nnz)
        if nbytes < 0:
            raise _lib.LpfError("graph too large for the 32-bit unit index of the packed link rows")
        # (torch's caching allocator aligns to 512 bytes; the rows need 128)
        self.slab = torch.empty(max(lib.lpf_link_rows_slab_bytes(n) // 4, 32), dtype=torch.int32, device=dev)
        self.overflow = torch.empty(max(nbytes // 4, 32), dtype=torch.int32, device=dev)
        scratch = torch.empty(max(lib.lpf_link_rows_scratch_bytes(n) // 8 + 2, 2), dtype=torch.int64, device=dev)
        call("lpf_pack_link_rows", ptr(adj.rowptr), ptr(adj.col), ptr(ppr.rowptr), ptr(ppr.col), ptr(ppr.val), n,
             ptr(self.slab), ptr(self.overflow), ptr(scratch), stream())
        torch.cuda.current_stream().synchronize()     # scratch is released here
        self.adj, self.ppr = adj, ppr


def link_rows(adj: CSR, ppr: CSR) -> LinkRows:
    """The packed rows of (adj, ppr), built once per table pair and cached on the adjacency object."""
    cache = adj.__dict__.setdefault("_link_rows", {})
    key = id(ppr)
    lr = cache.get(key)
    if lr is None or lr.ppr is not ppr:
        lr = cache[key] = LinkRows(adj, ppr)
    return lr


def select_onepass(links, adj: CSR, ppr: CSR, th_cn, th_1hop, th_non1hop, mode: str, cap: int, algo=None):
    """K1, one launch sequence and no host round trip (lpf_select_onepass): pairs of type t land in rows
    [t*cap, t*cap + header[t]) of the pair arrays, per-link segments in (seg_start, counts).  Returns a dict of the
    raw device buffers; ScorePlan (plan.py) is the production caller, tests use this wrapper."""
    require_cuda(links, adj.rowptr, ppr.rowptr)
    dev = links.device
    bs = links.shape[1]
    algo = pick_select_algo(adj, ppr, th_1hop, th_non1hop, mode) if algo is None else algo
    out = {
        "counts": torch.empty(3 * bs, dtype=torch.int32, device=dev),
        "seg_start": torch.empty(3 * bs, dtype=torch.int32, device=dev),
        "nz": torch.empty(bs, dtype=torch.int32, device=dev),
        "header": torch.zeros(8, dtype=torch.int64, device=dev),
        "node": torch.empty(3 * cap, dtype=torch.int32, device=dev),
        "src_ppr": torch.empty(3 * cap, dtype=torch.float32, device=dev),
        "tgt_ppr": torch.empty(3 * cap, dtype=torch.float32, device=dev),
        "cap": cap,
    }
    ws = torch.empty(_lib.load().lpf_select_workspace_bytes(bs) // 4, dtype=torch.int32, device=dev)
    out["workspace"] = ws        # (word 0: deferred links; word bs + 4: candidates of the packed screening)
    if algo == _lib.ALGO_PACKED:
        lr = link_rows(adj, ppr)
        call("lpf_select_onepass_packed", ptr(links), bs, ptr(adj.rowptr), ptr(adj.col), ptr(ppr.rowptr), ptr(ppr.col),
             ptr(ppr.val), ptr(lr.slab), ptr(lr.overflow), float(th_cn), float(th_1hop), float(th_non1hop), MODE[mode], cap,
             ptr(out["counts"]), ptr(out["seg_start"]), ptr(out["nz"]), ptr(out["header"]), ptr(out["node"]),
             ptr(out["src_ppr"]), ptr(out["tgt_ppr"]), ptr(ws), stream(), meta=(bs,))
        return out
    call("lpf_select_onepass", ptr(links), bs, ptr(adj.rowptr), ptr(adj.col), ptr(ppr.rowptr), ptr(ppr.col), ptr(ppr.val),
         float(th_cn), float(th_1hop), float(th_non1hop), MODE[mode], algo, cap, ptr(out["counts"]),
         ptr(out["seg_start"]), ptr(out["nz"]), ptr(out["header"]), ptr(out["node"]), ptr(out["src_ppr"]),
         ptr(out["tgt_ppr"]), ptr(ws), stream(), meta=(bs,))
    return out


def rpe_hidden(sel: Selection, t: int, w1, b1, ln_w, ln_b, hsum):
    r0, r1 = sel.type_range(t)
    if r1 > r0:
        call("lpf_rpe_hidden", ptr(sel.src_ppr), ptr(sel.tgt_ppr), r0, r1 - r0, ptr(w1), ptr(b1), ptr(ln_w),
             ptr(ln_b), hsum.shape[1], ptr(hsum), hsum.stride(0), None, stream())


ATTEND_WS_BYTES = 4 << 20     # chunk records of the grid-wide split of giant links (lpf_attend_fused_ws)


def attend(sel: Selection, KV, R, Q, att, bias, ln_w, ln_b, heads, ch, write_counts, out, alpha_out=None, idx=None,
           r_map=None, r_const=None):
    """K4 over every link of the batch, or over the batch positions in idx (rows of Q / out follow idx).  With r_map
    (int32 [S]) R holds only the rows of the pairs the map points to; the others read r_const [3, H*C] by type."""
    require_cuda(KV, R, Q, out, idx, r_map, r_const)
    n = sel.bs if idx is None else idx.numel()
    # links with thousands of pairs are split over the grid through a workspace (lpf_attend_fused_ws)
    ws = torch.empty(ATTEND_WS_BYTES, dtype=torch.uint8, device=KV.device)
    call("lpf_attend_fused_ws", ptr(sel.ptr), sel.bs, ptr(idx), n, ptr(sel.node), ptr(KV), KV.stride(0),
         ptr(R) if sel.total > 0 else None, R.stride(0) if R is not None and R.dim() == 2 else heads * ch,
         ptr(Q), Q.stride(0), ptr(att), ptr(bias), ptr(ln_w), ptr(ln_b), heads, ch, MODE[sel.mode],
         int(write_counts), ptr(out), out.stride(0), ptr(alpha_out), None, None, None, 0,
         int(KV.dtype == torch.bfloat16), ptr(r_map), ptr(r_const), ptr(ws), ws.numel(), stream(),
         meta=(n, sel.total, heads * ch))
    return out


def link_heads(links, X, consts, prob, idx=None, zb=None, logits=False, sched=None):
    """Fused tensor-core heads (lpf_link_heads_f16 / lpf_link_heads_tc): prob[pos] for every link (constant pairwise half `c3`) or for the
    positions in idx with per-row zb.  `sched` (int32 [2], zeros): tiles handed out dynamically (see the header)."""
    require_cuda(links, X, prob, idx, zb, sched)
    X = _rowmajor(X)
    bs = links.shape[1]
    n = bs if idx is None else idx.numel()
    heads_call(links, bs, ptr(idx), n, X, consts, ptr(zb), zb.stride(0) if zb is not None else 0, prob, logits, None,
               ptr(sched), stream())
    return prob


def gcn_layer(adj: CSR, XW, bias, ln=None, relu=False, residual=None, ln2=None, out=None, row0=0, rows=None, local=False):
    """One GCN layer after its Linear, in one launch (lpf_gcn_layer): out = LN2(residual + act(LN(A_hat XW + bias)))
    for rows [row0, row0 + rows).  ln / ln2: (weight, bias) or None.  local=False: `out` and `residual` are [n, d]
    tables indexed by node id; local=True: they hold only the rows of the range (row 0 = node row0: the row-sharded
    multi-GPU layers).  Falls back to lpf_gcn_spmm + lpf_layernorm_act for widths / alignments the fused kernel does
    not take."""
    require_cuda(adj.rowptr, XW, bias, out, residual)
    XW = _rowmajor(XW)
    n, d = adj.n, XW.shape[1]
    rows = n - row0 if rows is None else rows
    if out is None:
        out = torch.empty((rows if local else n, d), dtype=torch.float32, device=XW.device)
    if residual is not None:
        residual = _rowmajor(residual)
    shift = row0 if local else 0          # the kernel indexes `out` / `residual` by node id
    a16 = lambda t: t is None or t.data_ptr() % 16 == 0     # noqa: E731
    fusable = d % 4 == 0 and d <= 512 and XW.stride(0) % 4 == 0 and out.stride(0) % 4 == 0 and a16(XW) and a16(out) and \
        a16(bias) and (residual is None or (a16(residual) and residual.stride(0) % 4 == 0)) and \
        all(a16(t) for pair in (ln, ln2) if pair is not None for t in pair)
    if fusable:
        call("lpf_gcn_layer", ptr(adj.rowptr), ptr(adj.col), ptr(adj.val), row0, rows, ptr(XW), XW.stride(0), ptr(bias), d,
             ptr(ln[0]) if ln else None, ptr(ln[1]) if ln else None, int(relu),
             (residual.data_ptr() - 4 * shift * residual.stride(0)) if residual is not None else None,
             residual.stride(0) if residual is not None else 0,
             ptr(ln2[0]) if ln2 else None, ptr(ln2[1]) if ln2 else None,
             out.data_ptr() - 4 * shift * out.stride(0), out.stride(0), stream(), meta=(rows, d, adj.nnz))
        return out
    view = out if not local else None
    if local:       # unfused kernels index by node id: go through a full-height scratch table
        view = torch.empty((n, d), dtype=torch.float32, device=XW.device)
    gcn_spmm(adj, XW, bias, out=view, row0=row0, rows=rows)
    sh = view[row0:row0 + rows]
    dst = out if local else out[row0:row0 + rows]
    res = None if residual is None else (residual if local else residual[row0:row0 + rows])
    if ln is not None or relu or res is not None:
        layernorm_act(sh, None if ln is None else ln[0], None if ln is None else ln[1], relu=relu, residual=res, out=dst)
        sh = dst
    if ln2 is not None:
        layernorm_act(sh, ln2[0], ln2[1], relu=False, out=dst)
    elif sh is not dst:
        dst.copy_(sh)
    return out


def gcn_spmm(adj: CSR, XW, bias, out=None, row0=0, rows=None):
    require_cuda(adj.rowptr, XW, bias, out)
    XW = _rowmajor(XW)
    n, d = adj.n, XW.shape[1]
    rows = n - row0 if rows is None else rows
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=XW.device)
    call("lpf_gcn_spmm", ptr(adj.rowptr), ptr(adj.col), ptr(adj.val), row0, rows, ptr(XW), XW.stride(0), ptr(bias), d,
         ptr(out), out.stride(0), stream())
    return out
