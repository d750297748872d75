#!/usr/bin/env python
"""bench.py — scored links/s of the per-link pairwise-encoding path (BASELINE.json metric).

A "step" = one pass of the hot path (selection -> RPE -> attention -> heads -> mlp_score) over one
batch of synthetic queries on a precomputed X_node, i.e. the body of the reference's eval loop.
Default workload (BASELINE.json's headline config): the ogbl-citation2-shaped synthetic graph (2.93M
nodes, 30.6M edges, dim 64, hyper-parameters of reference scripts/replicate_heart.sh:22), 2,048 queries
of 1 held-out positive + 1,000 random negatives sharing the source (reference train/testing.py:14-47)
per step (= one batch of the eval driver: 2,050,048 links).  --workload {collab, ddi, ppa, cora} runs the other BASELINE shapes with HeaRT-style queries
(1 positive + 500 negatives, half random / half 2-hop corruptions of the target; train/testing.py:95-121)
in steps of the script's test batch size (scripts/replicate_heart.sh:4-19).

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path, one JSON line
  python bench.py --impl reference [...]                   the reference's CPU algorithm (oracle port)

Under torchrun (N > 1) queries are sharded over the ranks (weak scaling: --queries per rank per
step), node tables are replicated by one NCCL all-gather before the timed region, and the timed
region has no collective.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="citation2", choices=["citation2", "ppa", "collab", "ddi", "cora"])
    ap.add_argument("--scale", type=float, default=1.0, help="graph size multiplier (1.0 = the named shape)")
    ap.add_argument("--queries", type=int, default=None, help="queries per step per GPU (x (1+negs) links); default: "
                    "2048 for citation2, the script's test batch size // (1 + negs) for the other workloads")
    ap.add_argument("--negs", type=int, default=None)
    ap.add_argument("--cpu-sample-links", type=int, default=32768)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sampler", action="store_true", help="debug: do not poll nvidia-smi during the timed region")
    ap.add_argument("--depth", type=int, default=4, help="execution plans (batches) in flight in the pipelined eval loop")
    ap.add_argument("--node-dtype", default="f32", choices=["f32", "bf16"], help="storage of the node tables X / KV read by "
                    "the scoring plans (arithmetic is fp32 / fp16-split with fp32 accumulation either way)")
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=100):
        """Started BEFORE the warm-up (nvidia-smi's own start-up takes tens of ms and would otherwise land inside a
        short timed region); mark() / stop() bracket the timed region and only samples between them are used."""
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu=timestamp,{self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", str(period_ms), "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def mark(self):
        self.t0 = time.time()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.t1 = time.time()
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        import datetime
        parsed = []
        for r in rows:
            if len(r) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                parsed.append((ts, float(r[2]), float(r[3]), [v.strip().lower() == "active" for v in r[6:10]]))
            except ValueError:
                continue
        inside = [q for q in parsed if self.t0 is not None and self.t0 - 0.02 <= q[0] <= self.t1 + 0.12]
        if not inside and parsed:       # region shorter than the sampling period: the sample nearest to it
            mid = 0.5 * ((self.t0 or self.t1) + self.t1)
            inside = [min(parsed, key=lambda q: abs(q[0] - mid))]
        reasons = set()
        for q in inside:
            for name, on in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), q[3]):
                if on:
                    reasons.add(name)
        if inside:
            out.update(sm_mhz=float(np.median([q[1] for q in inside])), sm_max_mhz=float(max(q[2] for q in inside)),
                       reasons=sorted(reasons), samples=len(inside))
        return out


def make_workload(args, rank):
    from lpformer_b200 import synthetic as S
    cfg = S.CONFIGS[args.workload]
    negs = cfg["negs"] if args.negs is None else args.negs
    t0 = time.time()
    g = S.make_graph(args.workload, seed=args.seed, scale=args.scale, heldout=8192)
    gen_s = time.time() - t0
    return g, negs, gen_s


def queries_per_step(args, cfg, negs):
    if args.queries is not None:
        return args.queries
    # citation2: 2,048 queries = 2,050,048 links per batch.  The batch size is a parameter of the eval driver
    # (train/testing.py: `batch_size`), not of the workload; measured on a B200: 140 / 130 / 121 / 114 us per 256,256
    # links at 256 / 512 / 1,024 / 2,048 queries per batch (launch, pipeline-fill and staging costs of the nine
    # kernels of a batch are per launch)
    # (round 2, final kernels: 2.48 G links/s at 1,024 queries per batch, 2.69 G at 2,048 — the default; pipeline depth 4 /
    # 6 / 8 makes no difference)
    return 2048 if args.workload == "citation2" else max(1, cfg["batch"] // (1 + negs))


def make_queries(g, workload, nq, negs, seed):
    """[2, nq * (1 + negs)] int64 links of one step: citation2-style (shared source, uniform negatives) or HeaRT-style."""
    from lpformer_b200 import synthetic as S
    if workload == "citation2":
        return S.citation2_queries(g, nq, negs, seed=seed)
    return S.heart_queries(g, nq, negs, seed=seed)


def workload_name(workload, negs):
    style = "1 positive + %d negatives per query" % negs
    if workload != "citation2":
        style = "HeaRT-style eval, " + style + " (half random, half 2-hop corruptions)"
    else:
        style = "eval, " + style
    return f"ogbl-{workload}-shaped synthetic {style}" if workload != "cora" else f"Cora-shaped synthetic {style}"


def batch_bytes(g, links, negs, S_total, d, hc, nz_mask=None):
    """Algorithmic bytes of one batch: per link 4(deg a+deg b) [adj col ids] +
    8(nP a + nP b) [PPR col+val] + 32 [4 rowptr pairs] + 16 [link ids]; `select` is what one selection
    pass reads, dedup'd = the shared source's rows counted once per query instead of once per link."""
    deg = np.diff(g.indptr)
    npp = np.diff(g.ppr[0])
    a, b = links[0], links[1]
    row_a = 4 * deg[a] + 8 * npp[a]
    row_b = 4 * deg[b] + 8 * npp[b]
    per_link_fixed = 32 + 16
    select_full = float(row_a.sum() + row_b.sum() + per_link_fixed * links.shape[1])
    q = links.shape[1] // (1 + negs)
    row_a_dedup = row_a.reshape(q, 1 + negs)[:, 0].sum() if q * (1 + negs) == links.shape[1] else row_a.sum()
    select_dedup = float(row_a_dedup + row_b.sum() + (32 / 2 + 8) * links.shape[1] + 24 * q)
    e = 4
    path_extra = float(links.shape[1] * (2 * d * e + 4) + S_total * hc * e)
    out = {"select_full": select_full, "select_dedup": select_dedup, "path_full": select_full + path_extra,
           "path_dedup": select_dedup + path_extra - float((links.shape[1] - q) * d * e)}
    if nz_mask is not None:
        # the links that select something: the resolution reads both of their rows again (count -> allocate -> write),
        # the non-empty-link stage their two node rows, per selected pair a K/V row, the pair's RPE row (written once,
        # read once) and (node, ppr a, ppr b), and writes a score
        nnz = int(nz_mask.sum())
        out["resolve"] = float((row_a[nz_mask] + row_b[nz_mask]).sum() + per_link_fixed * nnz + 12 * S_total)
        out["nz_stage"] = float(nnz * (2 * d * e + 16 + 4) + S_total * (hc * e + 2 * d * e + 12))
    return out


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle/ref_port.py, torch sparse-COO algebra on the
    host cores) on a bounded sample of the same workload per step."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    from oracle import ref_port as R
    import lpformer_b200 as L
    from lpformer_b200 import synthetic as S
    torch.set_num_threads(os.cpu_count())
    g, negs, gen_s = make_workload(args, 0)
    cfg = g.cfg
    targs = S.train_args_of(cfg)
    torch.manual_seed(args.seed)
    model = L.LinkTransformer(targs, {"x": torch.from_numpy(g.x)}, device="cpu").eval()   # seeded weights only
    score = L.mlp_score(model.out_dim, model.out_dim, 1, 2).eval()
    P = {k: v.detach() for k, v in model.state_dict().items()}
    Sd = {k: v.detach() for k, v in score.state_dict().items()}
    X = torch.randn(g.n, cfg["dim"], generator=torch.Generator().manual_seed(args.seed + 5))
    A = R.coo_from_csr(g.indptr, g.indices, None, g.n)
    Pm = R.coo_from_csr(*g.ppr, g.n)
    nq_step = queries_per_step(args, cfg, negs)
    nq = max(1, min(nq_step, args.cpu_sample_links // (1 + negs)))
    cfgd = dict(targs)
    times = []
    for step in range(args.warmup + args.steps):
        # the SAME seeded batch the CUDA arm scores in this step (rank 0), of which the first `nq` queries are timed: the
        # reference's sparse index_select is O(nnz) per call, so a whole 2,048-query step would take minutes of CPU time
        full = make_queries(g, args.workload, nq_step, negs, seed=1000 + step)
        links = torch.from_numpy(np.ascontiguousarray(full[:, : nq * (1 + negs)]))
        t0 = time.perf_counter()
        R.score_links(links, X, A, Pm, P, Sd, cfgd, model.mask)
        if step >= args.warmup:
            times.append(time.perf_counter() - t0)
    nlinks = nq * (1 + negs)
    total = sum(times)
    val = nlinks * len(times) / total
    line = {"impl": "reference", "metric": "scored links/sec", "value": val, "unit": "links/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, negs), "scale": args.scale, "graph": g.stats(),
                       "queries_per_step_per_gpu": nq_step, "links_per_step_per_gpu": nq_step * (1 + negs),
                       "dim": cfg["dim"], "mode": model.mask,
                       "thresholds": [cfg["thresh_cn"], cfg["thresh_1hop"], cfg["thresh_non1hop"]]},
            "cpu_baseline": {"value": val, "unit": "links/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"the first {nq} of the {nq_step} queries of every step ({nlinks} links; the same "
                                       f"seeded batches as the CUDA arm) through oracle/ref_port.py — the reference's torch "
                                       f"sparse-COO algorithm, pinned to the unmodified reference by tests/test_reference_live.py; "
                                       f"the unmodified reference itself needs /root/reference, which does not exist on the GPU "
                                       f"box; X_node seeded N(0,1)"},
            "e2e": {"value": val, "unit": "links/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import lpformer_b200 as L
    from lpformer_b200 import _lib, synthetic as S
    from lpformer_b200.evaluate import propagate_replicated

    rank, local_rank, world = env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lpformer_b200 has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    g, negs, gen_s = make_workload(args, rank)
    # started here so that nvidia-smi's own start-up is long over when the timed region begins
    sampler = ClockSampler(local_rank) if (rank == 0 and not args.no_sampler) else None
    cfg = g.cfg
    targs = S.train_args_of(cfg)
    torch.manual_seed(args.seed)
    model = L.LinkTransformer(targs, g.data_dict(dev), device=dev).to(dev).eval()
    model.node_dtype = args.node_dtype
    score = L.mlp_score(model.out_dim, model.out_dim, 1, 2).to(dev).eval()
    d, hc = cfg["dim"], cfg["dim"]

    # ---- per-eval work (not in the timed region): GCN + gnn_norm + K/V projection (+ one all-gather)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    X = propagate_replicated(model)          # first call builds the normalised CSR
    torch.cuda.synchronize()
    e0.record()
    X = propagate_replicated(model)
    e1.record()
    torch.cuda.synchronize()
    propagate_ms = e0.elapsed_time(e1)
    # per-call breakdown of the per-eval work on rank 0 (eager, CUDA events around every C-ABI call): the GCN layer
    # launches are gathers of 4d-byte rows — algorithmic bytes per layer nnz (8 + 4d) + rows (8d + 8)
    per_eval_calls, gcn_roof = [], None
    if rank == 0 and world == 1:
        tr = _lib.Trace(events=True)
        _lib.TRACE = tr
        propagate_replicated(model)
        torch.cuda.synchronize()
        _lib.TRACE = None
        agg = {}
        for name, meta, a, b in tr.records:
            c, t = agg.get(name, (0, 0.0))
            agg[name] = (c + 1, t + a.elapsed_time(b))
        per_eval_calls = [{"name": k, "calls": c, "total_ms": t} for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
        lay = [(meta, a.elapsed_time(b)) for name, meta, a, b in tr.records if name == "lpf_gcn_layer" and meta]
        if lay:
            byt_l = sum(m[2] * (8 + 4 * m[1]) + m[0] * (8 * m[1] + 8) for m, _ in lay)
            ms_l = sum(t for _, t in lay)
            gcn_roof = {"kernel": "lpf_gcn_layer (gcn_spmm_kernel: SpMM + bias + LayerNorm + ReLU + residual)", "bound": "hbm",
                        "launches": len(lay), "avg_launch_ms": ms_l / len(lay), "algorithmic_bytes_per_launch": byt_l / len(lay),
                        "achieved": byt_l / (ms_l * 1e-3) / 1e9, "unit": "GB/s", "peak": load_peaks()[0],
                        "frac": byt_l / (ms_l * 1e-3) / 1e9 / load_peaks()[0],
                        "note": "gathered rows are counted once per edge (no credit for L2 reuse of a hub's row)"}

    nq = queries_per_step(args, cfg, negs)
    total_steps = args.warmup + args.steps
    host_links = [make_queries(g, args.workload, nq, negs, seed=1000 + rank * 100003 + s) for s in range(total_steps)]
    dev_links = [torch.from_numpy(l).to(dev) for l in host_links]
    nlinks = nq * (1 + negs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): the K batches through the pipelined eval loop (evaluate.LinkScoreStream:
    # the reference's batch loop, train/testing.py:25-32, with two execution plans in flight), between two events
    from lpformer_b200.evaluate import LinkScoreStream
    scorer = LinkScoreStream(model, score, X, nlinks, depth=args.depth)
    W = max(args.warmup, 3 * args.depth)        # both plans of the stream capture their CUDA graphs during the warm-up
    warm_links = torch.cat([dev_links[s % max(1, args.warmup)] for s in range(W)], dim=1)
    timed_dev = torch.cat(dev_links[args.warmup:], dim=1)
    timed_host = torch.cat([torch.from_numpy(l) for l in host_links[args.warmup:]], dim=1).pin_memory()
    out_host = torch.empty(timed_host.shape[1], dtype=torch.float32).pin_memory()
    scorer.score(warm_links)
    for _ in range(3):
        # a plan whose pair pools overflowed during the warm-up (dense graphs: ogbl-ppa shape, 540 k pairs per batch) was
        # rebuilt with larger pools at the end of that call and has yet to run eagerly once and capture its CUDA graphs
        if scorer.plans is None or all(P.runs >= 3 for P in scorer.plans):
            break
        scorer.score(warm_links)
    prewarmed = scorer.plans is None
    if prewarmed:
        # host-sized path (no plans: d = 128 / 256 or a threshold of 0 on the 1-hop / >1-hop sets): every batch allocates
        # pair-sized buffers (gigabytes on the ogbl-ddi shape) through torch's caching allocator, and a batch larger than
        # any before costs cudaMallocs inside the timed region (measured: 21 vs 8 ms per step).  One untimed pass over the
        # timed batches brings the allocator to the steady state the >= 200-batch e2e loop measures in anyway.
        scorer.score(timed_dev)
    barrier()
    if sampler:
        sampler.mark()           # (started long ago; no idle gap here — the GPU would drop its clocks)
    _lib.COUNTERS = {}
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    out = scorer.score(timed_dev)
    t_end.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    dev_ms = t_start.elapsed_time(t_end)
    graph_launches = _lib.COUNTERS.get("graph_launches", 0)
    _lib.COUNTERS = None

    # ---- the same K steps again, one score_links call per batch with every C-ABI call bracketed by CUDA events
    # (per-kernel breakdown; eager launches, no graph)
    for s in range(args.warmup):
        model.score_links(dev_links[s], X, score)
    barrier()
    trace = _lib.Trace(events=True)
    _lib.TRACE = trace
    outs = []
    import ctypes
    lib = _lib.load()
    lib.lpf_debug_select_timing(1)        # CUDA events around the kernels inside the selection entry point
    sel_ms, ms3 = [0.0, 0.0, 0.0, 0.0, 0], (ctypes.c_float * 4)()
    nz_ms = [0.0, 0.0, 0]
    for s in range(args.warmup, total_steps):
        # a 1 ms spin kernel first: the host enqueues the whole step behind it, so that every event pair brackets its
        # kernel's run time and not the launch latency of a GPU that waits for the host
        torch.cuda._sleep(2_000_000)
        outs.append(model.score_links(dev_links[s], X, score))
        if lib.lpf_debug_select_timing_read(ctypes.addressof(ms3)) == 0:
            for k in range(4):
                sel_ms[k] += ms3[k]
            sel_ms[4] += 1
        if lib.lpf_debug_nz_timing_read(ctypes.addressof(ms3)) == 0:
            nz_ms[0] += ms3[0]
            nz_ms[1] += ms3[1]
            nz_ms[2] += 1
    lib.lpf_debug_select_timing(0)
    barrier()
    _lib.TRACE = None
    stream_vs_single = float((torch.cat(outs) - out).abs().max())

    # ---- end to end through the public API with HOST buffers: every step's links come from pinned host memory
    # and its scores go back to pinned host memory inside the timed region
    # (window: the K batches cycled until >= 200 batches have gone through, so that the wall-clock figure is stable)
    scorer.score(timed_host[:, :2 * nlinks], out_host=out_host[:2 * nlinks])
    e2e_reps = max(1, -(-200 // args.steps)) if nlinks * args.steps < (1 << 26) else 1
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_reps):
        scorer.score(timed_host, out_host=out_host)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - w0) / e2e_reps
    barrier()
    e2e_vs_dev = float((out_host.to(dev) - out).abs().max())

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = times.tolist()

    if rank == 0:
        summ = trace.summary()
        launches_per_step = trace.launches / args.steps
        # set statistics + algorithmic bytes of the timed batches
        sel_stats = {"pairs_per_link": 0.0, "empty_frac": 0.0}
        tot_pairs, tot_empty, byt = 0, 0, {"select_full": 0.0, "select_dedup": 0.0, "path_full": 0.0, "path_dedup": 0.0,
                                           "resolve": 0.0, "nz_stage": 0.0}
        per_type, max_type = np.zeros(3), np.zeros(3)
        for s in range(args.warmup, total_steps):
            sel = model._select(dev_links[s], False)
            cn = sel.counts()
            c = cn.sum(0)
            tot_pairs += int(c.sum())
            tot_empty += int((c == 0).sum())
            per_type += cn.sum(1).cpu().numpy()
            max_type = np.maximum(max_type, cn.max(1).values.cpu().numpy())
            bb = batch_bytes(g, host_links[s], negs, sel.total, d, hc, nz_mask=(c > 0).cpu().numpy())
            for k in byt:
                byt[k] += bb[k]
        sel_stats["pairs_per_link"] = tot_pairs / (nlinks * args.steps)
        sel_stats["empty_frac"] = tot_empty / (nlinks * args.steps)
        for t, nm in enumerate(("cn", "1hop", "non1hop")):
            sel_stats["mean_" + nm] = float(per_type[t] / (nlinks * args.steps))
            sel_stats["max_" + nm] = int(max_type[t])

        if sel_ms[4] and "lpf_select_onepass_packed" in summ:
            # the selection entry point is several kernels: time them separately (events recorded inside the launcher)
            c, t = summ.pop("lpf_select_onepass_packed")
            summ["lpf_select_onepass_packed/screen (select_screen_packed_kernel)"] = (sel_ms[4], sel_ms[0])
            summ["lpf_select_onepass_packed/resolve (select_resolve_packed_kernel)"] = (sel_ms[4], sel_ms[1])
            summ["lpf_select_onepass_packed/hub-hub resolve (select_resolve_big_kernel)"] = (sel_ms[4], sel_ms[2])
            summ["lpf_select_onepass_packed/deferred links + finalize (select_heavy_onepass_kernel, select_finalize_onepass)"] = (sel_ms[4], sel_ms[3])
            summ["lpf_select_onepass_packed/launch gaps + resets"] = (c, max(0.0, t - sum(sel_ms[:4])))
        if nz_ms[2] and "lpf_nz_links_fused" in summ:
            c, t = summ.pop("lpf_nz_links_fused")
            summ["lpf_nz_links_fused/pair stage (nz_pairs_kernel)"] = (nz_ms[2], nz_ms[0])
            summ["lpf_nz_links_fused/link stage (nz_fused_kernel)"] = (nz_ms[2], nz_ms[1])
            summ["lpf_nz_links_fused/launch gaps"] = (c, max(0.0, t - nz_ms[0] - nz_ms[1]))
        kern = sorted(((n, c, t) for n, (c, t) in summ.items()), key=lambda r: -r[2])
        top_name, top_calls, top_ms = [k for k in kern if "launch gaps" not in k[0]][0]
        peak, peak_src = load_peaks()
        peaks_all = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
        avg_ms = top_ms / top_calls
        grouped = args.workload == "citation2"
        alg = alg_full = alg_flops = None
        if top_name.startswith("lpf_select_onepass_packed/screen") or top_name in (
                "lpf_select_count", "lpf_select_fill", "lpf_select_onepass", "lpf_select_onepass_packed"):
            # per link 4 deg + 8 nP of both rows + row pointers + ids; dedup'd = the shared source once per query
            alg = (byt["select_dedup"] if grouped else byt["select_full"]) / top_calls
            alg_full = byt["select_full"] / top_calls
        elif top_name.startswith("lpf_select_onepass_packed/") and "resolve" in top_name:
            # (both rows of every selecting link; a latency-bound kernel: dependent reads per candidate, not a stream)
            alg = alg_full = byt["resolve"] / top_calls
        elif top_name.startswith("lpf_nz_links_fused"):
            # (a latency-bound fp32 FFMA kernel on ~1 % of the links: bytes per non-empty link and selected pair)
            alg = alg_full = byt["nz_stage"] / top_calls
        elif top_name in ("lpf_link_heads_tc", "lpf_link_heads_f16"):
            # per link: X[b] row + link ids + score; the query's shared X[a] row once per query (dedup'd)
            alg = (args.steps * (nlinks * (d * 4 + 16 + 4) + nq * d * 4)) / top_calls
            alg_full = args.steps * nlinks * (2 * d * 4 + 16 + 4) / top_calls
            alg_flops = args.steps * nlinks * 2.0 * (d * d + d * d + 2 * d * d + 2 * d) / top_calls
        else:
            # contractions / attention of the unfused path: bytes and flops from the shapes recorded with every call
            tot_b = tot_f = 0.0
            for name, meta, _, _ in trace.records:
                if name != top_name or meta is None:
                    continue
                if name in ("lpf_gemm_tc", "lpf_gemm") and len(meta) == 3:
                    M, N, K = meta
                    tot_b += 4.0 * (M * K + M * N + N * K)
                    tot_f += 2.0 * M * N * K
                elif name in ("lpf_attend_fused", "lpf_attend_fused_ws") and len(meta) == 3:
                    n_l, S_p, HC = meta          # per pair a K/V row and an RPE row, per link a query row and an output row
                    tot_b += 4.0 * HC * (2 * S_p + 2 * n_l) + 4.0 * S_p
                    tot_f += 8.0 * HC * S_p
            if tot_b > 0:
                alg = alg_full = tot_b / top_calls
                alg_flops = tot_f / top_calls if tot_f > 0 else None
        traffic = None
        tpath = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tpath):       # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures
            m = re.search(r"\((\w+)", top_name)                 # "entry point/part (kernel, ...)" -> kernel
            kname = m.group(1) if m else top_name.replace("lpf_", "") + "_kernel"
            traffic = json.load(open(tpath)).get(kname) if grouped else None
        # (a shape whose node / graph tables fit the 126 MB L2 — ogbl-ddi: 4.4 MB of K/V rows — serves its gathers from L2:
        # `achieved` then counts L2-served bytes and `frac` relates them to the HBM peak for reference only)
        roof = {"bound": "hbm", "kernel": top_name, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": avg_ms,
                "share_of_kernel_time": top_ms / sum(t for _, _, t in kern)}
        if alg is not None:
            roof["achieved"] = alg / (avg_ms * 1e-3) / 1e9
            roof["frac"] = roof["achieved"] / peak
            roof["achieved_undeduped"] = alg_full / (avg_ms * 1e-3) / 1e9
            roof["algorithmic_bytes_per_launch"] = alg
        if alg_flops is not None:
            tf = alg_flops / (avg_ms * 1e-3) / 1e12
            tpeak = float(peaks_all.get("bf16_tflops_sustained", 1400.0))
            roof["tensor"] = {"achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                              "note": "algorithmic fp32 flops (2MNK of the contractions), executed as 3 split-precision "
                                      "MMAs each (fp16 hi/lo halves on kind::f16 at d = 64, tf32 otherwise); peak = "
                                      "measured sustained bf16"}
        # the selection's screening kernel is always reported too (the north star's HBM-roofline kernel)
        screen = [k for k in kern if k[0].startswith("lpf_select_onepass_packed/screen")]
        if screen:
            sc_ms = screen[0][2] / screen[0][1]
            sc_alg = (byt["select_dedup"] if grouped else byt["select_full"]) / screen[0][1]
            roof["screen"] = {"kernel": screen[0][0], "avg_launch_ms": sc_ms, "algorithmic_bytes_per_launch": sc_alg,
                              "achieved": sc_alg / (sc_ms * 1e-3) / 1e9, "frac": sc_alg / (sc_ms * 1e-3) / 1e9 / peak,
                              "traffic": (json.load(open(tpath)).get("select_screen_packed_kernel") if os.path.exists(tpath) and grouped else None)}
        # ... and so is the fused heads kernel (HBM gather + tensor pipe)
        heads_k = [k for k in kern if k[0] in ("lpf_link_heads_tc", "lpf_link_heads_f16")]
        if heads_k and grouped:
            h_ms = heads_k[0][2] / heads_k[0][1]
            h_alg = (args.steps * (nlinks * (d * 4 + 16 + 4) + nq * d * 4)) / heads_k[0][1]
            h_fl = args.steps * nlinks * 2.0 * (d * d + d * d + 2 * d * d + 2 * d) / heads_k[0][1]
            tpeak = float(peaks_all.get("bf16_tflops_sustained", 1400.0))
            roof["heads"] = {"kernel": heads_k[0][0], "avg_launch_ms": h_ms, "algorithmic_bytes_per_launch": h_alg,
                             "achieved": h_alg / (h_ms * 1e-3) / 1e9, "frac": h_alg / (h_ms * 1e-3) / 1e9 / peak,
                             "tensor_tflops": h_fl / (h_ms * 1e-3) / 1e12, "tensor_frac": h_fl / (h_ms * 1e-3) / 1e12 / tpeak,
                             "traffic": (json.load(open(tpath)).get("link_heads_f16_kernel") if os.path.exists(tpath) else None)}
        path_gbs = byt["path_dedup"] / (dev_ms * 1e-3) / 1e9
        value = world * nlinks * args.steps / (dev_ms * 1e-3)
        e2e_val = world * nlinks * args.steps / (e2e_ms * 1e-3)

        cpu_base = None
        if not args.no_cpu_baseline:
            cpu_base = cpu_baseline(args, g, negs, model, score, X, targs)

        table_gb = (g.indices.nbytes + g.ppr[1].nbytes * 2 + 2 * g.n * d * 4) / 1e9
        l2_policy = "distinct link batch every step; node/graph tables (%.2f GB) %s the 126 MB L2" % (
            table_gb, "exceed" if table_gb > 0.126 else "FIT in (the per-batch pair buffers, %.2f GB, do not)"
            % (sel_stats["pairs_per_link"] * nlinks * 2 * d * 4 / 1e9))
        if prewarmed:
            l2_policy += "; host-sized path: one untimed pass over the timed batches first (allocator steady state)"
        line = {"metric": "scored links/sec", "value": value, "unit": "links/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.workload, negs),
                           "scale": args.scale, "graph": g.stats(), "queries_per_step_per_gpu": nq,
                           "links_per_step_per_gpu": nlinks, "dim": d, "mode": model.mask, "node_table_dtype": args.node_dtype,
                           "thresholds": [cfg["thresh_cn"], cfg["thresh_1hop"], cfg["thresh_non1hop"]],
                           "l2_policy": l2_policy,
                           "set_stats": sel_stats, "parallelism": f"links sharded over {world} GPU(s), tables replicated"},
                "e2e": {"value": e2e_val, "unit": "links/s", "h2d_bytes_per_step": int(host_links[0].nbytes),
                        "d2h_bytes_per_step": int(nlinks * 4), "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(trace.launches),
                "api": {"depth": args.depth, "value": "evaluate.LinkScoreStream.score(device links): `depth` plans in flight, 1 CUDA-graph launch per "
                                 "batch, overflow flags read one batch late",
                        "e2e": "evaluate.LinkScoreStream.score(pinned host links, out_host=pinned scores): H2D / D2H on a "
                               "copy stream inside the timed region; wall clock over %d passes of the K batches" % e2e_reps,
                        "fast_path": scorer.plans is not None,
                        "max_abs_diff_stream_vs_score_links": stream_vs_single,
                        "max_abs_diff_e2e_vs_device": e2e_vs_dev},
                "gpu_launches_per_step": launches_per_step,
                "cuda_graph_launches_per_step": graph_launches / args.steps,
                "roofline": roof,
                "path_algorithmic_gbs": path_gbs,
                "path_frac_of_peak": path_gbs / peak,
                "kernels": [{"name": n, "calls": c, "total_ms": t} for n, c, t in kern],
                "per_eval": {"propagate_ms": propagate_ms, "graph_gen_s": gen_s, "calls": per_eval_calls, "roofline": gcn_roof},
                "cpu_baseline": cpu_base,
                "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, g, negs, model, score, X, targs):
    """The reference's CPU algorithm (oracle/ref_port.py) on a bounded sample of the same workload, same
    weights and the same X_node, on this box's host cores; also cross-checks the GPU scores."""
    from oracle import ref_port as R
    from lpformer_b200 import synthetic as S
    torch.set_num_threads(os.cpu_count())
    P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    Sd = {k: v.detach().cpu() for k, v in score.state_dict().items()}
    Xc = X.detach().cpu().contiguous()
    A = R.coo_from_csr(g.indptr, g.indices, None, g.n)
    Pm = R.coo_from_csr(*g.ppr, g.n)
    nq = max(1, min(queries_per_step(args, g.cfg, negs), args.cpu_sample_links // (1 + negs)))
    spent, done, first, max_diff = 0.0, 0, None, None
    for it in range(4):
        links_np = make_queries(g, args.workload, nq, negs, seed=5000 + it)
        links = torch.from_numpy(links_np)
        t0 = time.perf_counter()
        prob, _ = R.score_links(links, Xc, A, Pm, P, Sd, dict(targs), model.mask)
        dt = time.perf_counter() - t0
        if it == 0:
            first = dt
            gpu = model.score_links(links.to(X.device), X, score).cpu()
            max_diff = float((gpu - prob).abs().max())
        else:
            spent += dt
            done += 1
        if spent + first > 20.0:
            break
    if done == 0:
        spent, done = first, 1
    return {"value": nq * (1 + negs) * done / spent, "unit": "links/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{done} timed batch(es) of {nq} queries x {1 + negs} links after 1 warm-up batch "
                      f"(oracle/ref_port.py: the reference's torch sparse-COO algorithm)",
            "max_abs_prob_diff_vs_gpu": max_diff}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
