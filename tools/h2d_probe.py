"""Probe: H2D bandwidth of pinned link batches (contiguous vs 2-row strided slice), on the GPU box."""
import time
import torch
dev = torch.device("cuda:0")
L = 256256 * 20
h = torch.randint(0, 1000, (2, L), dtype=torch.int64).pin_memory()
d = torch.empty((2, 4 * 256256), dtype=torch.int64, device=dev)
s = torch.cuda.Stream()
for name, fn in [("strided 2-row slice", lambda g: d.copy_(h[:, g:g + d.shape[1]], non_blocking=True)),
                 ("row by row", lambda g: (d[0].copy_(h[0, g:g + d.shape[1]], non_blocking=True),
                                           d[1].copy_(h[1, g:g + d.shape[1]], non_blocking=True)))]:
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for g in range(0, L, d.shape[1]):
                fn(g)
        s.synchronize()
        dt = time.perf_counter() - t0
    print("%-22s %.2f ms for %.1f MB -> %.1f GB/s" % (name, dt * 1e3, h.numel() * 8 / 1e6, h.numel() * 8 / dt / 1e9))
o = torch.empty(L, dtype=torch.float32, device=dev)
oh = torch.empty(L, dtype=torch.float32).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); oh.copy_(o, non_blocking=True); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("D2H %.2f ms for %.1f MB -> %.1f GB/s" % (dt * 1e3, L * 4 / 1e6, L * 4 / dt / 1e9))
