"""PPR precompute: GPU push (lpf_ppr_push) vs the host port of the reference's numba kernel on the box's cores, on a
BASELINE-shaped synthetic graph; checks the two tables bit for bit.  Run on the GPU box:
    python tools/ppr_bench.py [workload] [scale]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lpformer_b200 import ppr, synthetic as S  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "citation2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
g = S.make_graph(workload, seed=0, scale=scale, heldout=8192)
eps = g.cfg["eps"]
dev = torch.device("cuda:0")
t0 = time.perf_counter()
want = S.ppr_push(g.indptr, g.indices, 0.15, eps)
host_s = time.perf_counter() - t0
indptr, indices = torch.from_numpy(g.indptr).to(dev), torch.from_numpy(g.indices).to(dev)
ppr.ppr_push(indptr, indices, 0.15, eps)          # warm-up (allocations, pool sizing)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
got = ppr.ppr_push(indptr, indices, 0.15, eps, cap=int(want[1].size) + 1024)
e1.record()
torch.cuda.synchronize()
gpu_s = e0.elapsed_time(e1) * 1e-3
same = (np.array_equal(got.rowptr.cpu().numpy(), want[0]) and np.array_equal(got.col.cpu().numpy(), want[1]) and
        np.array_equal(got.val.cpu().numpy().view(np.uint32), want[2].view(np.uint32)))
print(json.dumps({"workload": f"ogbl-{workload}-shaped synthetic, scale {scale}", "nodes": g.n, "edges": int(g.indices.size // 2),
                  "eps": eps, "ppr_nnz": int(want[1].size), "gpu_s": gpu_s, "gpu_sources_per_s": g.n / gpu_s,
                  "host_s": host_s, "host_cores": os.cpu_count(), "host_sources_per_s": g.n / host_s,
                  "bit_identical": bool(same)}))
