"""Runs the packed one-pass selection a few times on a citation2-shaped batch (target for ncu):
python tools/prof_select.py [scale] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lpformer_b200 import _lib, ops, synthetic as S  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
g = S.make_graph("citation2", seed=0, scale=scale, heldout=8192)
cfg = g.cfg
d = g.data_dict(dev)
th = (cfg["thresh_cn"], cfg["thresh_1hop"], cfg["thresh_non1hop"])
lib = _lib.load()
import ctypes
ms3 = (ctypes.c_float * 4)()
lib.lpf_debug_select_timing(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(iters):
    links = torch.from_numpy(S.citation2_queries(g, 256, 1000, seed=1000 + it)).to(dev)
    flush.zero_()
    out = ops.select_onepass(links, d["adj_mask"], d["ppr"], *th, "all", cap=1 << 19, algo=3)
    torch.cuda.synchronize()
    lib.lpf_debug_select_timing_read(ctypes.addressof(ms3))
    print("screen %.1f us  resolve %.1f us  hub-hub resolve %.1f us  deferred %.1f us" % (1e3 * ms3[0], 1e3 * ms3[1], 1e3 * ms3[2], 1e3 * ms3[3]), out["header"].tolist()[:5],
          "candidates %d deferred %d" % (int(out["workspace"][links.shape[1] + 4]), int(out["workspace"][0])))
