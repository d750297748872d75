"""A few eager score_links steps on the citation2-shaped workload (target for ncu of the heads / non-empty-link kernels):
python tools/prof_step.py [steps] [queries per step]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpformer_b200 as L  # noqa: E402
from lpformer_b200 import synthetic as S  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
g = S.make_graph("citation2", seed=0, scale=1.0, heldout=8192)
torch.manual_seed(0)
model = L.LinkTransformer(S.train_args_of(g.cfg), g.data_dict(dev), device=dev).to(dev).eval()
score = L.mlp_score(model.out_dim, model.out_dim, 1, 2).to(dev).eval()
X = torch.randn(g.n, g.cfg["dim"], device=dev)
for s in range(steps):
    links = torch.from_numpy(S.citation2_queries(g, nq, 1000, seed=1000 + s)).to(dev)
    out = model.score_links(links, X, score)
    torch.cuda.synchronize()
print("ok", float(out.mean()))
