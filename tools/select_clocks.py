"""Per-phase clock totals of the packed-row screening kernel (lpf_select_onepass_packed) on a citation2-shaped batch, and
the phase breakdown of its slowest piece.  Run on the GPU box:  python tools/select_clocks.py [scale]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lpformer_b200 import _lib, ops, synthetic as S  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
algo = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
g = S.make_graph("citation2", seed=0, scale=scale, heldout=8192)
cfg = g.cfg
d = g.data_dict(dev)
th = (cfg["thresh_cn"], cfg["thresh_1hop"], cfg["thresh_non1hop"])
lib = _lib.load()
buf = torch.zeros(64 + 3 * 2048, dtype=torch.int64, device=dev)
for it in range(6):
    links = torch.from_numpy(S.citation2_queries(g, 256, 1000, seed=1000 + it)).to(dev)
    if it == 5:
        lib.lpf_debug_select_clocks(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.select_onepass(links, d["adj_mask"], d["ppr"], *th, "all", cap=1 << 19, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    print("select_onepass %.1f us" % (1e3 * e0.elapsed_time(e1)), out["header"].tolist()[:5])
lib.lpf_debug_select_clocks(None)
t = buf.cpu().numpy()
n = max(1, int(t[5]))
print("pieces %d" % n)
ph = {10: "links + slab loads issued", 12: "source slab lines", 0: "source overflow + filters", 1: "screen (warp 0)", 2: "push"}
for k, nm in ph.items():
    print("  %-28s %9.0f cycles/piece  (%.2f us)" % (nm, t[k] / n, t[k] / n / 1965.0))

# per-piece timeline (globaltimer, ns): duration histogram and pieces in flight over time
pc = t[64:64 + 3 * n].reshape(-1, 3)
st, en, smid = pc[:, 0], pc[:, 1], pc[:, 2]
t00 = st.min()
dur = (en - st) / 1e3
print("piece duration us: min %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f; kernel span %.1f us" % (
    dur.min(), np.percentile(dur, 10), np.percentile(dur, 50), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), (en.max() - t00) / 1e3))
print("start times us: p50 %.1f p90 %.1f max %.1f" % tuple((np.percentile(st - t00, q) / 1e3) for q in (50, 90, 100)))
for tt in range(0, int((en.max() - t00) / 1e3) + 1, 4):
    live = int(((st - t00) / 1e3 <= tt).sum() - ((en - t00) / 1e3 <= tt).sum())
    print("  t=%3d us: %4d pieces in flight" % (tt, live))
slow = np.argsort(-dur)[:8]
deg = np.diff(g.indptr)
lk = S.citation2_queries(g, 256, 1000, seed=1000 + 5)
for q in slow:
    a = lk[0, q * 256:(q + 1) * 256]
    b = lk[1, q * 256:(q + 1) * 256]
    print("  slow piece %4d: %.1f us, start %.1f us, SM %d, source deg %s, max target deg %d" % (q, dur[q], (st[q] - t00) / 1e3, smid[q], sorted(set(deg[a].tolist())), deg[b].max()))
