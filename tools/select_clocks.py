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
buf = torch.zeros(48, dtype=torch.int64, device=dev)
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
names = ["flatten fill + stage source", "phase A", "phase B", "phase C", "generic"]
print("chunks %d, links resolved by a warp per piece %.1f, by the whole CTA per piece %.1f" % (n, t[6] / n, t[7] / n))
print("slowest chunk %.1f us, slowest CTA %.1f us" % (t[8] / 1965.0, t[9] / 1965.0))
for nm, v in zip(["  top loads+zero stores", "  run detection", "  table layout", "  clear+flatten scan"], t[10:14]):
    print("  %-26s %9.0f cycles/chunk  (%.1f us)" % (nm, v / n, v / n / 1965.0))
for nm, v in zip(names, t[:5]):
    print("  %-14s %9.0f cycles/chunk  (%.1f us)" % (nm, v / n, v / n / 1965.0))

ph = {10: "top loads", 12: "table layout", 13: "clear+flatten", 0: "stage source", 1: "phase A", 2: "phase B", 3: "phase C (CTA walks)", 4: "generic"}
print("slowest chunk: len %d runs %d items %d slow %d cta-walked %d table slots %d hub-launch %d" % tuple(int(x) for x in t[32:39]))
for k, nm in ph.items():
    print("    %-22s %8d cycles (%.1f us)" % (nm, t[16 + k], t[16 + k] / 1965.0))
