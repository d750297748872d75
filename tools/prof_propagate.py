"""Per-call breakdown of LinkTransformer.propagate() on the citation2-shaped graph: python tools/prof_propagate.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpformer_b200 as L  # noqa: E402
from lpformer_b200 import _lib, synthetic as S  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "citation2"
dev = torch.device("cuda:0")
g = S.make_graph(name, seed=0, scale=1.0, heldout=8192)
torch.manual_seed(0)
model = L.LinkTransformer(S.train_args_of(g.cfg), g.data_dict(dev), device=dev).to(dev).eval()
for _ in range(3):
    X = model.propagate()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    X = model.propagate()
e1.record()
torch.cuda.synchronize()
print("propagate %.2f ms" % (e0.elapsed_time(e1) / 5))
tr = _lib.Trace(events=True)
_lib.TRACE = tr
X = model.propagate()
torch.cuda.synchronize()
_lib.TRACE = None
for name_, meta, a, b in tr.records:
    print("  %-22s %8.3f ms  %s" % (name_, a.elapsed_time(b), meta))
adj = model.get_adj(False)
print("nnz %d  n %d  dim %d" % (adj.nnz, adj.n, g.cfg["dim"]))
