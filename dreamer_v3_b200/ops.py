"""Isolated device ops of libdv3 (the same kernels the fused path uses), taking CUDA tensors.
Used by the parity tests and by utils/symlog.py, utils/two_hot.py."""
import ctypes as C

import torch

from . import _lib


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(x):
    if not (torch.is_tensor(x) and x.is_cuda):
        raise RuntimeError("dreamer_v3_b200.ops needs CUDA tensors; there is no CPU fallback")
    return x.contiguous().to(torch.float32)


def _ok(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc})")


def symlog(x):
    x = _f32(x)
    y = torch.empty_like(x)
    _ok(_lib.load().dv3_op_symlog(_p(x), _p(y), x.numel(), _s()), "dv3_op_symlog")
    return y


def inverse_symlog(y):
    y = _f32(y)
    x = torch.empty_like(y)
    _ok(_lib.load().dv3_op_inverse_symlog(_p(y), _p(x), y.numel(), _s()), "dv3_op_inverse_symlog")
    return x


def two_hot(value, dense=True):
    """255 buckets in [-20, 20]. Returns (k, kp1, w_k, w_kp1, dense [n,255] or None)."""
    v = _f32(value).reshape(-1)
    n = v.numel()
    k = torch.empty(n, dtype=torch.int32, device=v.device)
    kp1 = torch.empty_like(k)
    wk = torch.empty_like(v)
    wkp1 = torch.empty_like(v)
    d = torch.empty((n, 255), dtype=torch.float32, device=v.device) if dense else None
    _ok(_lib.load().dv3_op_two_hot(_p(v), n, _p(k), _p(kp1), _p(wk), _p(wkp1), _p(d), _s()), "dv3_op_two_hot")
    return k, kp1, wk, wkp1, d


def sample_categorical(probs, u):
    p = _f32(probs)
    K = p.shape[-1]
    rows = p.numel() // K
    uu = _f32(u).reshape(-1)
    assert uu.numel() == rows
    idx = torch.empty(rows, dtype=torch.int32, device=p.device)
    _ok(_lib.load().dv3_op_sample_categorical(_p(p), _p(uu), rows, K, _p(idx), _s()), "dv3_op_sample_categorical")
    return idx.reshape(p.shape[:-1])


def value_targets(r, c, v, gamma, lambda_):
    r, c, v = _f32(r), _f32(c), _f32(v)
    Hp1, N = r.shape
    out = torch.empty((Hp1 - 1, N), dtype=torch.float32, device=r.device)
    _ok(_lib.load().dv3_op_value_targets(_p(r), _p(c), _p(v), Hp1 - 1, N, gamma, lambda_, _p(out), _s()),
        "dv3_op_value_targets")
    return out


def percentiles(x):
    x = _f32(x).reshape(-1)
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    _ok(_lib.load().dv3_op_percentiles(_p(x), x.numel(), _p(out), _s()), "dv3_op_percentiles")
    return out


def gemm(A, B, ta=False, tb=False, bias=None, C_=None, beta=0.0, mode=0):
    """C = op(A) op(B) (+bias) (+ beta*C). A, B 2-D row-major CUDA tensors (may be column slices)."""
    assert A.dim() == 2 and B.dim() == 2 and A.stride(1) == 1 and B.stride(1) == 1
    M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    N = B.shape[0] if tb else B.shape[1]
    assert (B.shape[1] if tb else B.shape[0]) == K
    if C_ is None:
        C_ = torch.zeros((M, N), dtype=torch.float32, device=A.device)
    _ok(_lib.load().dv3_op_gemm(_p(A), A.stride(0), int(ta), _p(B), B.stride(0), int(tb), _p(C_), C_.stride(0), M, N, K,
                                _p(bias), beta, mode, _s()), "dv3_op_gemm")
    return C_


def gemm_weight_shadow(A, W, ta=False, bias=None, C_=None, beta=0.0):
    """C = op(A) W with W [K, N] given in fp32; its bf16 K-major shadow (what the engine keeps per weight matrix and
    refreshes once per optimizer step) is built here, outside any timing loop the caller runs around `run`."""
    assert A.dim() == 2 and W.dim() == 2 and A.stride(1) == 1
    Wh = W.t().contiguous().to(torch.bfloat16)  # [N, K]
    M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    N = W.shape[1]
    assert W.shape[0] == K and K % 8 == 0
    if C_ is None:
        C_ = torch.zeros((M, N), dtype=torch.float32, device=A.device)

    def run():
        _ok(_lib.load().dv3_op_gemm_shadow(_p(A), A.stride(0), int(ta), _p(Wh), K, _p(C_), C_.stride(0), M, N, K,
                                           _p(bias), beta, _s()), "dv3_op_gemm_shadow")
        return C_
    run.keepalive = Wh
    return run


def ln_silu_bwd(dy, x, gamma, beta, dgamma=None, dbeta=None):
    """dx of y = silu(LayerNorm(x)) (row statistics recomputed here with torch for the probe); accumulates dgamma / dbeta."""
    x = _f32(x)
    mean = x.mean(1).contiguous()
    rstd = torch.rsqrt(x.var(1, unbiased=False) + 1e-3).contiguous()
    dx = torch.empty_like(x)

    def run():
        _ok(_lib.load().dv3_op_ln_silu_bwd(_p(_f32(dy)), _p(x), _p(mean), _p(rstd), _p(_f32(gamma)), _p(_f32(beta)), _p(dx), _p(dgamma),
                                           _p(dbeta), x.shape[0], x.shape[1], _s()), "dv3_op_ln_silu_bwd")
        return dx
    return run


def to_bf16(x):
    """bf16 copy of a 2-D fp32 CUDA tensor, made by the library's conversion kernel."""
    x = _f32(x)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _ok(_lib.load().dv3_op_to_bf16(_p(x), x.stride(0), _p(out), out.stride(0), x.shape[0], x.shape[1], _s()), "dv3_op_to_bf16")
    return out


def gemm_bf16(A16, W, bias=None, C_=None, beta=0.0, small=False):
    """C = A16 W with A16 [M, K] bf16 (as written by the producer kernels in tensor-core mode) and W [K, N] fp32 whose
    bf16 K-major shadow is built here; `run` launches the TMA-fed tcgen05 kernel only (small=True: the small-shape
    mma.sync kernel of the dream rollout's per-step contractions)."""
    assert A16.dtype == torch.bfloat16 and A16.dim() == 2 and A16.stride(1) == 1
    Wh = W.t().contiguous().to(torch.bfloat16)  # [N, K]
    M, K = A16.shape
    N = W.shape[1]
    assert W.shape[0] == K and K % 8 == 0
    if C_ is None:
        C_ = torch.zeros((M, N), dtype=torch.float32, device=A16.device)

    def run():
        fn = _lib.load().dv3_op_gemm_bf16_small if small else _lib.load().dv3_op_gemm_bf16
        _ok(fn(_p(A16), A16.stride(0), _p(Wh), K, _p(C_), C_.stride(0), M, N, K, _p(bias), beta, _s()),
            "dv3_op_gemm_bf16_small" if small else "dv3_op_gemm_bf16")
        return C_
    run.keepalive = Wh
    return run


def gemm_bf16_wgrad(X16, dY16, C_=None, beta=0.0):
    """C [M, N] (+)= X16^T dY16 with X16 [K, M], dY16 [K, N] bf16 row-major: the weight-gradient contraction on the
    operands as stored (TMA boxes of 64 k x 64 mn, MN-major UMMA descriptors)."""
    assert X16.dtype == torch.bfloat16 and dY16.dtype == torch.bfloat16 and X16.shape[0] == dY16.shape[0]
    K, M = X16.shape
    N = dY16.shape[1]
    if C_ is None:
        C_ = torch.zeros((M, N), dtype=torch.float32, device=X16.device)
    _ok(_lib.load().dv3_op_gemm_bf16_wgrad(_p(X16), X16.stride(0), _p(dY16), dY16.stride(0), _p(C_), C_.stride(0), M, N, K,
                                           beta, _s()), "dv3_op_gemm_bf16_wgrad")
    return C_


def ln_silu(x, gamma, beta):
    x = _f32(x)
    y = torch.empty_like(x)
    _ok(_lib.load().dv3_op_ln_silu(_p(x), _p(_f32(gamma)), _p(_f32(beta)), _p(y), x.shape[0], x.shape[1], _s()),
        "dv3_op_ln_silu")
    return y
