"""GPU timeline of the pipelined eval loop (CUPTI through torch.profiler): which kernels run when, on which stream, and how
much of the time the GPU runs 1 / 2 / 3+ kernels at once.  python tools/timeline.py [batches] [depth]"""
import collections
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpformer_b200 as L  # noqa: E402
from lpformer_b200 import synthetic as S  # noqa: E402
from lpformer_b200.evaluate import LinkScoreStream  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
g = S.make_graph("citation2", seed=0, scale=1.0, heldout=8192)
torch.manual_seed(0)
model = L.LinkTransformer(S.train_args_of(g.cfg), g.data_dict(dev), device=dev).to(dev).eval()
score = L.mlp_score(model.out_dim, model.out_dim, 1, 2).to(dev).eval()
X = torch.randn(g.n, g.cfg["dim"], device=dev)
links = torch.cat([torch.from_numpy(S.citation2_queries(g, 256, 1000, seed=1000 + s)) for s in range(nb)], dim=1).to(dev)
bs = 256 * 1001
stream = LinkScoreStream(model, score, X, bs, depth=depth)
for _ in range(3):
    stream.score(links)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stream.score(links)
    e1.record()
    torch.cuda.synchronize()
print("%.1f us per batch (events)" % (1e3 * e0.elapsed_time(e1) / nb))
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
short = lambda n: n.split("(")[0].split("::")[-1].split("<")[0][:28]      # noqa: E731
print("%d kernels; first 40 of the steady state (start us, dur us, stream, name):" % len(ev))
mid = len(ev) // 2
for e in ev[mid:mid + 40]:
    print("  %9.1f %7.1f  s%-3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), short(e["name"])))
# concurrency histogram over the steady state
lo, hi = ev[len(ev) // 4]["ts"], ev[3 * len(ev) // 4]["ts"]
pts = []
for e in ev:
    pts.append((e["ts"], 1))
    pts.append((e["ts"] + e["dur"], -1))
pts.sort()
cur, last, hist = 0, None, collections.Counter()
for t, d in pts:
    if last is not None and lo <= last and t <= hi:
        hist[cur] += t - last
    cur += d
    last = t
tot = sum(hist.values())
print("kernels running at once (share of the steady-state time):", {k: round(v / tot, 3) for k, v in sorted(hist.items())})
busy = collections.Counter()
for e in ev:
    if lo <= e["ts"] <= hi:
        busy[short(e["name"])] += e["dur"]
nbat = sum(1 for e in ev if lo <= e["ts"] <= hi and "link_heads" in e["name"])
print("kernel time per batch (us):", {k: round(v / max(nbat, 1), 1) for k, v in busy.most_common()})
