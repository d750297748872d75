"""Sync-free execution plan of the eval-loop body for one batch size.

`LinkTransformer.score_links` (the body of reference train/testing.py:29-31 / :113-115) is a fixed sequence of
C-ABI launches whose row counts (selected pairs per type, links with a non-empty set) are produced ON the
device by the one-pass selection kernel.  A ScorePlan owns capacity-sized buffers for one batch size, passes the
device-side counts to every later launch (`*_dev` arguments of include/lpformer_b200.h), and records the whole
sequence once in a CUDA graph; a batch is then one H2D/D2D copy of the links, one graph launch and one
read-back of the overflow flag.  If a pair pool overflows the plan reports it and the caller re-runs the batch
through the host-sized two-pass path (and the plan is rebuilt with larger pools).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import EPI_NONE, MODE, call, ptr, stream


class ScorePlan:
    def __init__(self, model, score_func, consts, X_node, kv, bs, test_set, logits, cap=None, use_graph=True):
        dev = X_node.device
        self.model, self.bs, self.logits, self.dev = model, bs, bool(logits), dev
        self.score_func, self.test_set = score_func, test_set
        # node tables as the kernels read them: fp32, or bf16 copies (model.node_dtype == "bf16": half the gather bytes,
        # the arithmetic stays fp32 / fp16-split with fp32 accumulation)
        self.X, self.kv = model._node_tables(X_node, kv) if "w1h" in consts else (X_node, kv)
        self.tab_bf16 = int(self.X.dtype == torch.bfloat16)
        self.consts = consts
        self.adj = model.get_adj(test_set, mask=True)
        self.ppr = model.get_ppr(test_set)
        from . import ops
        self.algo = ops.pick_select_algo(self.adj, self.ppr, model.thresh_1hop, model.thresh_non1hop, model.mask)
        if self.algo == _lib.ALGO_GENERIC:
            raise _lib.LpfError("the one-pass plan needs an INTERSECT selection algorithm")
        # sparse graphs: thread-per-link screening over the packed link rows (built once per table pair)
        self.rows = ops.link_rows(self.adj, self.ppr) if (self.algo == _lib.ALGO_INTERSECT8 and ops.USE_PACKED_ROWS) else None
        d, H = model.dim, model.num_heads
        layer = model.att_layers[0]
        self.layer = layer
        self.C = layer.att.out_channels
        HC = H * self.C
        self.HC, self.d, self.H = HC, d, H
        self.cap = int(cap if cap is not None else max(1 << 16, 2 * bs))
        cap, f32, i32, i64 = self.cap, torch.float32, torch.int32, torch.int64
        e = lambda *shape, dtype=f32: torch.empty(shape, dtype=dtype, device=dev)   # noqa: E731
        self.links = e(2, bs, dtype=i64)
        self.prob = e(bs)
        self.counts, self.seg_start, self.nz = e(3 * bs, dtype=i32), e(3 * bs, dtype=i32), e(bs, dtype=i32)
        self.hdr = torch.zeros(8, dtype=i64, device=dev)
        self.sched = torch.zeros(2, dtype=i32, device=dev)      # tile scheduler words of lpf_link_heads_tc (self re-arming)
        # (measured on the citation2 shape: no gain over the fixed stride, 173 vs 167 us per step, so off by default)
        self.dynamic_tiles = False
        self.ws = e(_lib.load().lpf_select_workspace_bytes(bs) // 4, dtype=i32)
        self.node, self.pa, self.pb = e(3 * cap, dtype=i32), e(3 * cap), e(3 * cap)
        self.hsum, self.R = e(3 * cap, d), e(3 * cap, HC)
        pd = HC + model.count_dim
        self.pd = pd
        self.xsum, self.Q, self.feats, self.hid, self.pw, self.zb = e(bs, d), e(bs, HC), e(bs, pd), e(bs, pd), e(bs, d), e(bs, 2 * d)
        self.att_ws = torch.empty(4 << 20, dtype=torch.uint8, device=dev)    # lpf_attend_fused_ws: giant links' chunk records
        derived = model._get_derived()[0]
        self.rpe = []
        for enc, (m, c) in zip(model._encoders(), derived["rpe"]):
            self.rpe.append((enc.linears[0].weight.detach(), enc.linears[0].bias.detach(), enc.norm.weight.detach(),
                             enc.norm.bias.detach(), ops.pack_weight(m), c))
        att = layer.att
        pl = model.pairwise_lin
        self.w = {
            "wl": ops.pack_weight(att.lin_l.weight), "bl": att.lin_l.bias.detach(),
            "att": att.att.detach(), "abias": att.bias.detach(),
            "pn_w": layer.post_att_norm.weight.detach(), "pn_b": layer.post_att_norm.bias.detach(),
            "p1": ops.pack_weight(pl.linears[0].weight), "pb1": pl.linears[0].bias.detach(),
            "pln_w": pl.norm.weight.detach(), "pln_b": pl.norm.bias.detach(),
            "p2": ops.pack_weight(pl.linears[1].weight), "pb2": pl.linears[1].bias.detach(),
            "wz": ops.pack_weight(consts["ws1_pw"]),
        }
        # no non-linearity between pairwise_lin's last Linear and mlp_score's first: folded once (fp64) into one
        # contraction  zb = (Ws1[:, d:] W_p2) hid + (Ws1[:, d:] b_p2 + off)
        wz64 = consts["ws1_pw"].detach().double()
        self.wz2 = (wz64 @ pl.linears[1].weight.detach().double()).float().contiguous()
        self.w["wz2"] = ops.pack_weight(self.wz2)
        self.w["off2"] = (wz64 @ pl.linears[1].bias.detach().double() + consts["off"].double()).float().contiguous()
        self.th = (float(model.thresh_cn), float(model.thresh_1hop), float(model.thresh_non1hop))
        self.mode = MODE[model.mask]
        # small-batch fused path for the non-empty links (lpf_nz_links_fused): transposed fp32 weights
        T = lambda w: w.detach().float().t().contiguous()   # noqa: E731
        self.keep = [T(att.lin_l.weight), T(pl.linears[0].weight), T(pl.linears[1].weight), T(consts["ws1_pw"]),
                     T(model.elementwise_lin.linears[0].weight), T(consts["w23"])] + [T(m) for m, _ in derived["rpe"]]
        a = _lib.NzArgs()
        a.links, a.bs, a.nz, a.n_cap, a.n_dev = ptr(self.links), bs, ptr(self.nz), bs, self.hdr.data_ptr() + 3 * 8
        a.X, a.ldx, a.KV, a.ld_kv = ptr(self.X), self.X.stride(0), ptr(self.kv), self.kv.stride(0)
        a.tab_bf16 = self.tab_bf16
        a.node, a.src_ppr, a.tgt_ppr = ptr(self.node), ptr(self.pa), ptr(self.pb)
        a.seg_start, a.counts, a.cap, a.d, a.mode = ptr(self.seg_start), ptr(self.counts), cap, d, self.mode
        a.header, a.R = self.hdr.data_ptr(), ptr(self.R)
        a.wlT, a.bl = ptr(self.keep[0]), ptr(self.w["bl"])
        for t, (w1, b1, g, b, _, cvec) in enumerate(self.rpe):
            a.rpe_w1[t], a.rpe_b1[t], a.rpe_ln_w[t], a.rpe_ln_b[t] = ptr(w1), ptr(b1), ptr(g), ptr(b)
            a.rpe_mT[t], a.rpe_c[t] = ptr(self.keep[6 + t]), ptr(cvec)
        a.att, a.att_bias, a.post_ln_w, a.post_ln_b = ptr(self.w["att"]), ptr(self.w["abias"]), ptr(self.w["pn_w"]), ptr(self.w["pn_b"])
        a.p1T, a.pb1, a.pln_w, a.pln_b = ptr(self.keep[1]), ptr(self.w["pb1"]), ptr(self.w["pln_w"]), ptr(self.w["pln_b"])
        a.p2T, a.pb2, a.wzT, a.off = ptr(self.keep[2]), ptr(self.w["pb2"]), ptr(self.keep[3]), ptr(consts["off"])
        a.w1T, a.b1, a.ln_w, a.ln_b = ptr(self.keep[4]), ptr(consts["b1"]), ptr(consts["ln_w"]), ptr(consts["ln_b"])
        a.w23T, a.ws2, a.bs2, a.prob, a.logits = ptr(self.keep[5]), ptr(consts["ws2"]), ptr(consts["bs2"]), ptr(self.prob), int(self.logits)
        self.nz_args = a
        self.fused_ok = H == 1 and HC == d and d in (32, 64)
        self.nz_mode = "batched"        # switched to "fused" when the observed share of non-empty links is small
        self.graphs = {}
        # the tensor-core heads of ALL links (empty-set constant) do not depend on the selection: they run on a side
        # stream next to it (a fork / join inside the captured graph) and fill the SMs the selection's tail leaves idle
        self.side = torch.cuda.Stream(device=dev)
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        self.fused_allowed = True
        self.graph = None
        self.hdr_host = torch.zeros(8, dtype=torch.int64).pin_memory()
        self.done = torch.cuda.Event()
        self.last = [0] * 8
        self.use_graph = use_graph
        self.runs = 0

    # ------------------------------------------------------------------
    def _launch(self):
        """Every launch of one batch on the current stream; no host synchronisation, no allocation."""
        from . import ops
        st = stream()
        bs, cap, d, HC, pd = self.bs, self.cap, self.d, self.HC, self.pd
        c, w, hdr = self.consts, self.w, self.hdr
        links, X = self.links, self.X
        hp = hdr.data_ptr()
        n_dev = hp + 3 * 8                      # &header[3]: links with a non-empty set

        def heads(idx, n, zb, ndev, on=None):
            ops.heads_call(links, bs, idx, n, X, c, ptr(zb), self.zb.stride(0) if zb is not None else 0, self.prob,
                           self.logits, ndev, ptr(self.sched) if self.dynamic_tiles else None, st if on is None else on)

        def gemm(A, Wp, bias, scale, C, M, N, K, mdev):
            call("lpf_gemm_tc", ptr(A), A.stride(0), ptr(Wp), ptr(bias), float(scale), ptr(C), C.stride(0), M, N, K,
                 EPI_NONE, mdev, st, meta=(M, N, K))

        # every link with the empty-set pairwise constant (independent of the selection): side stream — except under
        # the per-kernel tracer (bench.py), where a kernel running next to the one being timed would inflate its time
        main = torch.cuda.current_stream()
        traced = _lib.TRACE is not None
        if traced:
            heads(None, bs, None, None)
        else:
            self.ev_fork.record(main)
            self.side.wait_event(self.ev_fork)
            with torch.cuda.stream(self.side):
                heads(None, bs, None, None, on=self.side.cuda_stream)
                self.ev_join.record(self.side)
        # K1: one-pass selection into the per-type pools
        if self.rows is not None:
            call("lpf_select_onepass_packed", ptr(links), bs, ptr(self.adj.rowptr), ptr(self.adj.col),
                 ptr(self.ppr.rowptr), ptr(self.ppr.col), ptr(self.ppr.val), ptr(self.rows.slab), ptr(self.rows.overflow),
                 *self.th, self.mode, cap, ptr(self.counts), ptr(self.seg_start), ptr(self.nz), hp, ptr(self.node),
                 ptr(self.pa), ptr(self.pb), ptr(self.ws), st, meta=(bs,))
        else:
            call("lpf_select_onepass", ptr(links), bs, ptr(self.adj.rowptr), ptr(self.adj.col), ptr(self.ppr.rowptr),
                 ptr(self.ppr.col), ptr(self.ppr.val), *self.th, self.mode, self.algo, cap, ptr(self.counts),
                 ptr(self.seg_start), ptr(self.nz), hp, ptr(self.node), ptr(self.pa), ptr(self.pb), ptr(self.ws), st,
                 meta=(bs,))
        if not traced:
            main.wait_event(self.ev_join)
        if self.nz_mode == "fused":
            # few non-empty links: one warp per link, everything from node sets to score in one launch
            call("lpf_nz_links_fused", C.byref(self.nz_args), st, meta=(bs,))
            return
        # RPE hidden vectors and their contraction: every pair of every type pool in one launch (FFMA, matrices in shared
        # memory) when the fused kernels cover the configuration; else per type pool rpe_hidden + a tensor-core contraction
        if self.fused_ok and self.nz_mode != "batched_tc":
            call("lpf_nz_pairs", C.byref(self.nz_args), st, meta=(bs,))
        else:
            for t, (w1, b1, g, b, mp, cvec) in enumerate(self.rpe):
                call("lpf_rpe_hidden", ptr(self.pa), ptr(self.pb), t * cap, cap, ptr(w1), ptr(b1), ptr(g), ptr(b), d,
                     ptr(self.hsum), self.hsum.stride(0), hp + t * 8, st)
                gemm(self.hsum[t * cap:], mp, cvec, 1.0, self.R[t * cap:], cap, HC, d, hp + t * 8)
        # compacted non-empty links: query vectors, attention, pairwise_lin, offset of mlp_score's first layer
        call("lpf_gather_links", ptr(links), bs, ptr(self.nz), bs, ptr(X), X.stride(0), d, ptr(self.xsum),
             self.xsum.stride(0), None, 0, n_dev, self.tab_bf16, st)
        gemm(self.xsum, w["wl"], w["bl"], 2.0, self.Q, bs, HC, d, n_dev)
        call("lpf_attend_fused_ws", None, bs, ptr(self.nz), bs, ptr(self.node), ptr(self.kv), self.kv.stride(0),
             ptr(self.R), self.R.stride(0), ptr(self.Q), self.Q.stride(0), ptr(w["att"]), ptr(w["abias"]),
             ptr(w["pn_w"]), ptr(w["pn_b"]), self.H, self.C, self.mode, 1, ptr(self.feats), self.feats.stride(0), None,
             n_dev, ptr(self.seg_start), ptr(self.counts), cap, self.tab_bf16, None, None, ptr(self.att_ws), self.att_ws.numel(), st,
             meta=(bs, 0, HC))
        gemm(self.feats, w["p1"], w["pb1"], 1.0, self.hid, bs, pd, pd, n_dev)
        call("lpf_layernorm_act", ptr(self.hid), self.hid.stride(0), ptr(w["pln_w"]), ptr(w["pln_b"]), None, 0,
             ptr(self.hid), self.hid.stride(0), bs, pd, 1, n_dev, st)
        gemm(self.hid, w["wz2"], w["off2"], 1.0, self.zb, bs, 2 * d, pd, n_dev)
        heads(ptr(self.nz), bs, self.zb, n_dev)

    def grow(self, factor=4):
        """Larger pair pools after an overflow (buffers that depend on the capacity are re-allocated, graphs dropped)."""
        self.__init__(self.model, self.score_func, self.consts, self.X, self.kv, self.bs, self.test_set, self.logits,
                      cap=self.cap * factor, use_graph=self.use_graph)

    def submit(self, links):
        """Launches one batch on the current stream without any host synchronisation: copy of `links` (int64 [2, bs],
        host-pinned or device) into the plan, the launch sequence (a CUDA-graph replay after the first batch), and an
        asynchronous copy of the header to pinned host memory.  `collect()` later tells whether a pair pool overflowed;
        `self.prob` holds the scores in stream order."""
        if links is not None:          # None: the caller has already filled self.links (on a stream this one waits for)
            if links.is_cuda or links.is_contiguous():
                self.links.copy_(links, non_blocking=True)
            else:                      # strided host slice: two contiguous row copies instead of a slow pitched one
                self.links[0].copy_(links[0], non_blocking=True)
                self.links[1].copy_(links[1], non_blocking=True)
        tracing = _lib.TRACE is not None
        if self.use_graph and not tracing and self.runs >= 1:
            g = self.graphs.get(self.nz_mode)
            if g is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self.graphs[self.nz_mode] = g
            self.graph = g
            g.replay()
            if _lib.COUNTERS is not None:
                _lib.COUNTERS["graph_launches"] = _lib.COUNTERS.get("graph_launches", 0) + 1
        else:
            self._launch()
        self.runs += 1
        self.hdr_host.copy_(self.hdr, non_blocking=True)
        self.done.record()

    def collect(self):
        """Waits for the batch submitted last and returns its overflow flag (the batch's only host round trip)."""
        self.done.synchronize()
        h = self.hdr_host.tolist()
        self.last = h
        # regime for the NEXT batch (either path is exact; this only picks the cheaper one)
        if self.fused_ok and self.fused_allowed:
            # (one warp per link in FFMA while the links fit one wave or two; the tensor-core sequence — six small
            # contractions, attention, LayerNorm, a second heads launch — beyond: measured 0.460 vs 0.444 ms per step at
            # 13 k non-empty links, 140 vs 186 us at 3 k)
            self.nz_mode = "fused" if (h[3] < self.model.nz_fused_share * self.bs and
                                       h[3] <= self.model.nz_fused_max_links) else "batched"
            # hundreds of thousands of selected pairs (dense graphs: ogbl-ppa shape, 540 k per 32,565-link batch): the FFMA
            # pair stage streams its d x d matrix through shared memory once per pair (0.8 ns per pair); the hidden
            # vectors written out and contracted on the tensor cores cost less from here on
            if self.nz_mode == "batched" and h[0] + h[1] + h[2] > self.model.nz_pairs_tc_min:
                self.nz_mode = "batched_tc"
        return bool(h[4])

    def run(self, links):
        """Scores `links` (int64 [2, bs], host-pinned or device).  Returns (prob [bs] — a buffer owned by the plan,
        valid until the next run — and overflow: bool)."""
        self.submit(links)
        return self.prob, self.collect()

    def stats(self):
        h = self.hdr.tolist()
        return {"pairs": h[:3], "nonempty_links": h[3], "overflow": h[4], "cap": self.cap}
