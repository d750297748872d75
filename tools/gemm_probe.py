"""Device time of lpf_gemm_tc on the shapes of the BASELINE workloads (CUDA events, 5 launches after 2 warm-up):
python tools/gemm_probe.py [one]   ('one': a single launch of the first shape, for an ncu capture)"""
import sys
import torch
sys.path.insert(0, ".")
from lpformer_b200 import ops

SHAPES = [(2_330_000, 256, 256, "ddi RPE contraction"), (1_000_000, 128, 128, "collab-like"), (540_000, 64, 64, "ppa RPE"),
          (2_930_000, 64, 128, "citation2 GCN layer 0 (per eval)"), (13_000, 64, 64, "citation2 non-empty links")]
one = len(sys.argv) > 1 and sys.argv[1] == "one"
dev = torch.device("cuda:0")
for M, N, K, what in SHAPES[:1] if one else SHAPES:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    reps = 1 if one else 5
    for _ in range(0 if one else 2):
        ops.linear(A, W, b, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.linear(A, W, b, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 4.0 * (M * K + M * N) / 1e9
    tf = 2.0 * M * N * K / 1e12
    err = (out[:4096] - (A[:4096].double() @ W.double().t() + b.double()).float()).abs().max().item()
    print("%-36s M=%8d N=%3d K=%3d  %8.3f ms  %6.0f GB/s  %6.1f TFLOP/s (x3 executed)  tile %.1f us  max err %.2e" %
          (what, M, N, K, ms, gb / ms * 1e3, tf / ms * 1e3, ms * 1e3 / ((M + 127) // 128 / 148.0), err))
    del A, out
