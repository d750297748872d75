"""Per-phase clock64() profile of the fused heads kernel's consumer pipeline (CTA 0, first 8 tiles).
Run on the GPU box:  python tools/heads_clocks.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpformer_b200 as L  # noqa: E402
from lpformer_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
n, d = 400000, 64
targs = dict(dim=d, num_heads=1, trans_layers=1, gnn_layers=1, residual=False, layer_norm=True, relu=True,
             thresh_cn=0, thresh_1hop=1e-3, thresh_non1hop=1e-2)
data = {"x": torch.zeros(n, 4)}
torch.manual_seed(0)
model = L.LinkTransformer(targs, data, device=dev).to(dev).eval()
score = L.mlp_score(2 * d, 2 * d, 1, 2).to(dev).eval()
X = torch.randn(n, d, device=dev)
consts = model._head_consts(score, X)
bs = 148 * 128 * 10
links = torch.randint(0, n, (2, bs), device=dev)
links[0] = links[0, 0]
prob = torch.empty(bs, device=dev)
buf = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
lib = _lib.load()
for rep in range(3):
    ops.link_heads(links, X, consts, prob)
torch.cuda.synchronize()
lib.lpf_debug_heads_clocks(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.link_heads(links, X, consts, prob)
e1.record()
torch.cuda.synchronize()
lib.lpf_debug_heads_clocks(None)
t = buf.cpu().numpy().reshape(8, 16)[:, :6]
names = ["wait_mma3(prev)", "epi2(prev)", "wait_mma1", "epi1+sync", "issue_mma3"]
print("kernel %.1f us for %d tiles/CTA" % (1e3 * e0.elapsed_time(e1), bs // 128 // 148))
for i in range(1, 8):
    dt = np.diff(t[i])
    nxt = (" period=%d" % (t[i + 1, 0] - t[i, 0])) if i < 7 else ""
    print("tile %d: " % i + "  ".join("%s=%d" % (nm, v) for nm, v in zip(names, dt)) + nxt)
