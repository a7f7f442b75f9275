"""tcgen05 / TMEM bf16 GEMM (csrc/dv3_gemm_tc.cu) against an fp64 product of the bf16-rounded operands
(isolates layout / descriptor errors from the expected bf16 input rounding), all four transposition modes,
ragged M/N/K, bias and accumulate epilogues, split-K."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(M, N, K, ta, tb, seed, mode=2):
    from dreamer_v3_b200 import ops
    g = torch.Generator().manual_seed(seed)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    Ar = A.bfloat16().double()
    Br = B.bfloat16().double()
    want = (Ar.T if ta else Ar) @ (Br.T if tb else Br)
    got = ops.gemm(A.cuda(), B.cuda(), ta, tb, bias=bias.cuda(), mode=mode).cpu().double()
    tol = 2e-5 * np.sqrt(K) * 8
    assert float((got - (want + bias.double())).abs().max()) <= tol, (M, N, K, ta, tb, "bias")
    got = ops.gemm(A.cuda(), B.cuda(), ta, tb, C_=C0.cuda(), beta=1.0, mode=mode).cpu().double()
    assert float((got - (want + C0.double())).abs().max()) <= tol, (M, N, K, ta, tb, "beta")


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 128), (256, 128, 192), (1024, 256, 1280), (1000, 255, 257),
                                   (130, 24, 70), (4096, 512, 48), (16384, 256, 1281), (1280, 256, 16384), (384, 1024, 640)])
def test_gemm_tc_matches_bf16_reference(M, N, K):
    for ta in (False, True):
        for tb in (False, True):
            _case(M, N, K, ta, tb, seed=M + 3 * N + 7 * K + int(ta) + 2 * int(tb))


def test_gemm_tc_strided_operands():
    from dreamer_v3_b200 import ops
    g = torch.Generator().manual_seed(5)
    big = torch.randn(512, 1400, generator=g).cuda()
    W = torch.randn(1280, 256, generator=g).cuda()
    A = big[:, 3:1283]  # unaligned column slice, row stride 1400
    got = ops.gemm(A, W, mode=2).cpu().double()
    want = A.cpu().bfloat16().double() @ W.cpu().bfloat16().double()
    assert float((got - want).abs().max()) < 2e-5 * np.sqrt(1280) * 8


def test_gemm_tc_close_to_fp32():
    """bf16 tolerance statement for the tensor-core mode: |C_tc - C_fp32| <= 2^-8 * sqrt(K) * |A|_rms |B|_rms (3 sigma)."""
    from dreamer_v3_b200 import ops
    g = torch.Generator().manual_seed(6)
    A, B = torch.randn(2048, 1280, generator=g).cuda(), torch.randn(1280, 256, generator=g).cuda()
    c_tc = ops.gemm(A, B, mode=2)
    c_32 = ops.gemm(A, B, mode=0)
    err = float((c_tc - c_32).abs().max())
    assert err < 3 * 2 ** -8 * np.sqrt(1280) , err


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1024, 768, 256), (1000, 255, 264), (16384, 256, 1280), (130, 24, 72),
                                   (1024, 1024, 256), (4096, 512, 48), (300, 128, 1096)])
def test_gemm_tc_weight_shadow_variant(M, N, K):
    """Warp-specialised kernel with the bf16 K-major weight shadow (what the engine uses for every weight operand)."""
    from dreamer_v3_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    for ta in (False, True):
        A = torch.randn((K, M) if ta else (M, K), generator=g)
        W = torch.randn(K, N, generator=g)
        bias = torch.randn(N, generator=g)
        C0 = torch.randn(M, N, generator=g)
        Ar = A.bfloat16().double()
        want = (Ar.T if ta else Ar) @ W.bfloat16().double()
        tol = 2e-5 * np.sqrt(K) * 8
        got = ops.gemm_weight_shadow(A.cuda(), W.cuda(), ta=ta, bias=bias.cuda())().cpu().double()
        assert float((got - (want + bias.double())).abs().max()) <= tol, (M, N, K, ta)
        got = ops.gemm_weight_shadow(A.cuda(), W.cuda(), ta=ta, C_=C0.cuda(), beta=1.0)().cpu().double()
        assert float((got - (want + C0.double())).abs().max()) <= tol, (M, N, K, ta, "beta")


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1024, 768, 256), (1000, 252, 264), (16384, 256, 1280), (130, 24, 72),
                                   (1024, 1024, 256), (4096, 512, 48), (300, 128, 1096), (20000, 1024, 512), (1024, 256, 4096)])
def test_gemm_tma_variant(M, N, K):
    """TMA-fed persistent kernel (csrc/dv3_gemm_tma.cu): both operands bf16 K-major in global memory. Ragged M/N/K,
    bias + accumulate epilogues, split-K, several tiles per CTA (double-buffered TMEM accumulator)."""
    from dreamer_v3_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g)
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    A16 = ops.to_bf16(A.cuda())
    assert torch.equal(A16.cpu(), A.bfloat16())
    want = A.bfloat16().double() @ W.bfloat16().double()
    tol = 2e-5 * np.sqrt(K) * 8
    got = ops.gemm_bf16(A16, W.cuda(), bias=bias.cuda())().cpu().double()
    assert float((got - (want + bias.double())).abs().max()) <= tol, (M, N, K)
    got = ops.gemm_bf16(A16, W.cuda(), C_=C0.cuda(), beta=1.0)().cpu().double()
    assert float((got - (want + C0.double())).abs().max()) <= tol, (M, N, K, "beta")


@pytest.mark.parametrize("M,N,K", [(1024, 256, 256), (1024, 768, 256), (1024, 256, 1280), (1024, 1024, 256), (1024, 256, 1024),
                                   (1024, 256, 768), (1000, 250, 264), (64, 16, 32), (130, 24, 72), (4096, 320, 1000)])
def test_gemm_small_variant(M, N, K):
    """Small-shape mma.sync kernel (csrc/dv3_gemm_small.cu): the per-step contractions of the XS dream rollout and its BPTT.
    Ragged M / N / K (K % 8 == 0), bias, accumulate (also into a strided view, as the BPTT does into the [h, z] gradient)."""
    from dreamer_v3_b200 import ops
    g = torch.Generator().manual_seed(M * 3 + N * 5 + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g)
    bias = torch.randn(N, generator=g)
    A16 = ops.to_bf16(A.cuda())
    want = A.bfloat16().double() @ W.bfloat16().double()
    tol = 2e-5 * np.sqrt(K) * 8
    got = ops.gemm_bf16(A16, W.cuda(), bias=bias.cuda(), small=True)().cpu().double()
    assert float((got - (want + bias.double())).abs().max()) <= tol, (M, N, K)
    wide = torch.randn(M, N + 6, generator=g).cuda()      # accumulate into columns [2, 2 + N) of a wider matrix
    before = wide.clone()
    ops.gemm_bf16(A16, W.cuda(), C_=wide[:, 2:2 + N], beta=1.0, small=True)()
    assert float((wide[:, 2:2 + N].cpu().double() - (want + before[:, 2:2 + N].cpu().double())).abs().max()) <= tol, (M, N, K, "beta")
    assert torch.equal(wide[:, :2], before[:, :2]) and torch.equal(wide[:, 2 + N:], before[:, 2 + N:])


@pytest.mark.parametrize("M,N,K", [(128, 64, 512), (1280, 256, 15360), (256, 256, 16384), (1000, 248, 1000), (64, 768, 1024),
                                   (1280, 256, 1024), (520, 72, 4104)])
def test_gemm_tma_wgrad_variant(M, N, K):
    """Weight-gradient contraction X^T dY on bf16 operands as stored: MN-major UMMA descriptors, 64x64 TMA boxes,
    split-K with TMA reduce-add, accumulate epilogue. Ragged M / N / K."""
    from dreamer_v3_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    X = torch.randn(K, M, generator=g)
    dY = torch.randn(K, N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    want = X.bfloat16().double().T @ dY.bfloat16().double()
    tol = 2e-5 * np.sqrt(K) * 8
    got = ops.gemm_bf16_wgrad(ops.to_bf16(X.cuda()), ops.to_bf16(dY.cuda())).cpu().double()
    assert float((got - want).abs().max()) <= tol, (M, N, K)
    got = ops.gemm_bf16_wgrad(ops.to_bf16(X.cuda()), ops.to_bf16(dY.cuda()), C_=C0.cuda(), beta=1.0).cpu().double()
    assert float((got - (want + C0.double())).abs().max()) <= tol, (M, N, K, "beta")
