import sys, time, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import lpformer_b200 as L
from lpformer_b200 import synthetic as S, _lib
from lpformer_b200.evaluate import LinkScoreStream, propagate_replicated
dev = torch.device("cuda:0")
g = S.make_graph("citation2", seed=0, scale=1.0, heldout=8192)
targs = S.train_args_of(g.cfg)
torch.manual_seed(0)
model = L.LinkTransformer(targs, g.data_dict(dev), device=dev).to(dev).eval()
score = L.mlp_score(model.out_dim, model.out_dim, 1, 2).to(dev).eval()
X = propagate_replicated(model)
nq, negs = 256, 1000
links = [torch.from_numpy(S.citation2_queries(g, nq, negs, seed=1000 + s)).to(dev) for s in range(40)]
allb = torch.cat(links, 1)
scorer = LinkScoreStream(model, score, X, nq * (1 + negs), depth=4)
scorer.score(allb[:, : 12 * nq * 1001]); torch.cuda.synchronize()
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = scorer.score(allb)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("40 batches: host loop %.1f us/batch, gpu events %.1f us/batch, wall %.1f us/batch" % ((t1 - t0) / 40 * 1e6, e0.elapsed_time(e1) / 40 * 1e3, (t2 - t0) / 40 * 1e6))
# pure host cost: how long does the loop take when the GPU is not the limit?  use a tiny batch size
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); scorer.score(allb); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
