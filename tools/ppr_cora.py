import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from lpformer_b200 import ppr, synthetic as S
for scale in (0.25, 1.0):
    t0 = time.time(); g = S.make_graph("cora", seed=3, scale=scale, heldout=64); t1 = time.time()
    dev = torch.device("cuda:0")
    for rep in range(3):
        torch.cuda.synchronize(); t2 = time.time()
        got = ppr.ppr_push(torch.from_numpy(g.indptr).to(dev), torch.from_numpy(g.indices).to(dev), 0.15, g.cfg["eps"])
        torch.cuda.synchronize(); t3 = time.time()
        ok = (np.array_equal(got.rowptr.cpu().numpy(), g.ppr[0]) and np.array_equal(got.col.cpu().numpy(), g.ppr[1])
              and np.array_equal(got.val.cpu().numpy().view(np.uint32), g.ppr[2].view(np.uint32)))
        print("scale %.2f n %d: host graph+ppr %.1f s, gpu ppr %.2f s, nnz %d, bit-exact %s" % (scale, g.n, t1 - t0, t3 - t2, got.col.numel(), ok), flush=True)
