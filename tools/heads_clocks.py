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
n, d = 2927963, 64
targs = dict(dim=d, num_heads=1, trans_layers=1, gnn_layers=1, residual=False, layer_norm=True, relu=True,
             thresh_cn=0, thresh_1hop=1e-3, thresh_non1hop=1e-2)
data = {"x": torch.zeros(n, 4)}
torch.manual_seed(0)
model = L.LinkTransformer(targs, data, device=dev).to(dev).eval()
score = L.mlp_score(2 * d, 2 * d, 1, 2).to(dev).eval()
X = torch.randn(n, d, device=dev)
consts = model._head_consts(score, X)
bs = 256 * 1001
links = torch.randint(0, n, (2, bs), device=dev)
links[0] = torch.randint(0, n, (256,), device=dev).repeat_interleave(1001)     # citation2-style runs of one source
prob = torch.empty(bs, device=dev)
buf = torch.zeros(400 + 2 * 148, dtype=torch.int64, device=dev)
lib = _lib.load()
for rep in range(400):          # (long enough for the clocks to ramp up)
    ops.link_heads(links, X, consts, prob)
torch.cuda.synchronize()
(lib.lpf_debug_heads_f16_clocks if "w1h" in consts else lib.lpf_debug_heads_clocks)(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.link_heads(links, X, consts, prob)
e1.record()
torch.cuda.synchronize()
(lib.lpf_debug_heads_f16_clocks if "w1h" in consts else lib.lpf_debug_heads_clocks)(None)
raw = buf.cpu().numpy()
full = raw[:256].reshape(16, 16)
print("kernel %.1f us for %.1f tiles/CTA -> SM clock %.2f GHz" % (1e3 * e0.elapsed_time(e1), bs / 128 / 148, (raw[257] - raw[256]) / (1e3 * e0.elapsed_time(e1)) / 1e3))
span = raw[400:400 + 296].reshape(148, 2)
if span[0, 1] > span[0, 0]:
    print("CTA 0 lived %.1f us (globaltimer) -> SM clock %.2f GHz" % ((span[0, 1] - span[0, 0]) / 1e3, (raw[257] - raw[256]) / (span[0, 1] - span[0, 0])))
    t0 = span[:, 0].min()
    st, en = (span[:, 0] - t0) / 1e3, (span[:, 1] - t0) / 1e3
    print("all CTAs (us, globaltimer): start min/median/max %.1f/%.1f/%.1f  end min/median/max %.1f/%.1f/%.1f  life min/median/max %.1f/%.1f/%.1f"
          % (st.min(), np.median(st), st.max(), en.min(), np.median(en), en.max(), (en - st).min(), np.median(en - st), (en - st).max()))
print("CTA 0: kernel start -> first tile top %d cycles; tile tops (consumer) relative to start: %s; end %d" % (full[0, 0] - raw[256], [int(full[i, 0] - raw[256]) for i in range(10)], raw[257] - raw[256]))
t = full[:, :7]
names = ["wait_mma1", "epi1", "wait_mma3(prev)", "store_h+sync", "issue_mma3", "epi2(prev)"]
for i in range(1, 8):
    dt = np.diff(t[i])
    nxt = (" period=%d" % (t[i + 1, 0] - t[i, 0])) if i < 7 else ""
    print("tile %d: " % i + "  ".join("%s=%d" % (nm, v) for nm, v in zip(names, dt)) + nxt)

# the MMA thread: [0] loop top, [1] next operand tile ready, [2] contraction 1 of tile it+1 issued, [3] H ready,
# [4] contraction 2 of tile it issued (an issue blocks while the MMA queue is full, so these are roughly run times)
m = full[:, 8:13]
for i in range(1, 7):
    d = np.diff(m[i])
    print("mma thread tile %d: wait_a=%d issue_mma1=%d wait_h=%d issue_mma3=%d  period=%d" % (i, d[0], d[1], d[2], d[3], m[i + 1, 0] - m[i, 0]))

# producer thread 0: [13] loop top, [14] gathered rows have arrived, [15] operand tile stored (after waiting for the
# previous contraction 1)
for i in range(1, 7):
    if full[i, 7]:
        print("producer tile %d: issue+sync=%d wait_rows=%d store=%d period=%d" % (i, full[i, 7] - full[i, 13], full[i, 14] - full[i, 7], full[i, 15] - full[i, 14], full[i + 1, 13] - full[i, 13]))
    else:
        print("producer tile %d: gather=%d wait+store=%d period=%d" % (i, full[i, 14] - full[i, 13], full[i, 15] - full[i, 14], full[i + 1, 13] - full[i, 13]))

p2 = raw[272:272 + 128].reshape(16, 8)
if p2[1, 2]:
    # relative to the loop top: next source row requested + operand buffer free, rows read + multiplied, maxima exchanged,
    # tile stored, fence, end barrier passed
    for i in range(1, 7):
        t0 = full[i, 13]
        print("producer detail tile %d: " % i + " ".join("%d" % (v - t0) for v in p2[i][2:]))
