"""-m gpu: the tcgen05 (3xTF32, TMEM accumulator) contraction against a float64 reference of the same op."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K)
    (128, 64, 64), (1000, 64, 64), (257, 68, 68), (129, 1, 128), (5, 64, 2), (4096, 128, 128), (300, 256, 256),
    (100, 512, 512), (777, 64, 1433), (1, 16, 32), (3000, 80, 40),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm_tc_matches_fp64(M, N, K, epi):
    from lpformer_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g)
    ref = A.double() @ W.double().T + 2.0 * b.double()
    if epi == 1:
        ref = ref.clamp_min(0)
    elif epi == 2:
        ref = torch.sigmoid(ref)
    old = ops.GEMM_BACKEND
    ops.GEMM_BACKEND = "tc"
    try:
        out = ops.linear(A.to(dev), W.to(dev), b.to(dev), bias_scale=2.0, epilogue=epi)
        # strided views: A with a wider leading dimension, output into a column slice
        Abig = torch.zeros(M, K + 5, device=dev)
        Abig[:, :K] = A.to(dev)
        Cbig = torch.full((M, N + 3), -7.0, device=dev)
        ops.linear(Abig[:, :K], W.to(dev), b.to(dev), bias_scale=2.0, out=Cbig[:, 1:N + 1], epilogue=epi)
    finally:
        ops.GEMM_BACKEND = old
    torch.cuda.synchronize()
    scale = float(ref.abs().max()) + 1e-6
    err = float((out.cpu().double() - ref).abs().max()) / scale
    # fp32-level: the tensor core's fp32 accumulator truncates, so the error grows ~linearly with the number of
    # K-slices; plain single-pass TF32 would sit at ~5e-4 regardless of K
    tol = max(1e-5, 4e-8 * K)
    assert err < tol, f"max scaled error {err:.3g} (tol {tol:.3g})"
    assert torch.equal(Cbig[:, 1:N + 1], out)
    assert bool((Cbig[:, 0] == -7.0).all()) and bool((Cbig[:, N + 1:] == -7.0).all())


@pytest.mark.parametrize("K,N", [(32, 64), (2, 64), (32, 68), (64, 64)])
def test_gemm_tc_persistent_grid_with_device_row_count(K, N):
    """A device-side row count caps the grid (persistent CTAs walk several tiles): a single-k-block contraction
    (K <= 32) then uses both ring stages across tiles.  M is large enough that every CTA takes more than one tile."""
    from lpformer_b200 import _lib, ops
    from lpformer_b200._lib import call, ptr, stream
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(K * 131 + N)
    cap, m_real = 148 * 4 * 128 * 3 + 77, 148 * 4 * 128 * 2 + 1234      # grid capped at 592 CTAs: 2-3 tiles each
    A = torch.randn(cap, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    Wp = ops.pack_weight(W)
    C = torch.full((cap, N), -3.0, device=dev)
    m_dev = torch.tensor([m_real], dtype=torch.int64, device=dev)
    call("lpf_gemm_tc", ptr(A), A.stride(0), ptr(Wp), ptr(b), 1.0, ptr(C), C.stride(0), cap, N, K, _lib.EPI_NONE,
         m_dev.data_ptr(), stream())
    torch.cuda.synchronize()
    ref = A[:m_real].double() @ W.double().T + b.double()
    scale = float(ref.abs().max()) + 1e-6
    assert float((C[:m_real].double() - ref).abs().max()) / scale < 1e-5
    assert bool((C[m_real:] == -3.0).all())          # rows beyond the device-side count are untouched
