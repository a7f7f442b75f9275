"""In-graph cost of one dependent kernel node for the small (N = 1024 rows) shapes of a dreamed step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dreamer_v3_b200 import ops  # noqa: E402


def time_chain(fn, reps=60):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


N = 1024
x = torch.randn(N, 256, device="cuda")
gam, bet = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
print(f"ln_silu 1024x256                      : {time_chain(lambda: ops.ln_silu(x, gam, bet)):.2f} us / node")
for (M, Nn, K) in ((1024, 256, 256), (1024, 768, 256), (1024, 1024, 256), (1024, 256, 1024), (1024, 256, 1280), (1024, 256, 768)):
    A16 = ops.to_bf16(torch.randn(M, K, device="cuda"))
    W = torch.randn(K, Nn, device="cuda")
    C = torch.empty(M, Nn, device="cuda")
    run = ops.gemm_bf16(A16, W, C_=C)
    print(f"gemm_tma {M}x{Nn}x{K:<5d}              : {time_chain(run):.2f} us / node")
A = torch.randn(N, 256, device="cuda"); W1 = torch.randn(256, 1, device="cuda"); C1 = torch.empty(N, 1, device="cuda")
print(f"gemm_f32 1024x1x256 (actor head)      : {time_chain(lambda: ops.gemm(A, W1, C_=C1, mode=0)):.2f} us / node")
A2 = torch.randn(N, 1, device="cuda"); W2 = torch.randn(1, 256, device="cuda"); C2 = torch.zeros(N, 256, device="cuda")
print(f"gemm_f32 1024x256x1 (action part)     : {time_chain(lambda: ops.gemm(A2, W2, C_=C2, beta=1.0, mode=0)):.2f} us / node")
A3 = torch.randn(16384, 256, device="cuda"); C3 = torch.empty(16384, 1, device="cuda")
print(f"gemm_f32 16384x1x256 (continue head)  : {time_chain(lambda: ops.gemm(A3, W1, C_=C3, mode=0)):.2f} us / node")
for rows in (1024, 16384):
    xx = torch.randn(rows, 256, device="cuda"); dy = torch.randn(rows, 256, device="cuda")
    dg, db = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    print(f"ln_silu_bwd {rows}x256 (dx only)        : {time_chain(ops.ln_silu_bwd(dy, xx, gam, bet)):.2f} us / node")
    print(f"ln_silu_bwd {rows}x256 (+dgamma, dbeta) : {time_chain(ops.ln_silu_bwd(dy, xx, gam, bet, dg, db)):.2f} us / node")
    print(f"ln_silu     {rows}x256                  : {time_chain(lambda: ops.ln_silu(xx, gam, bet)):.2f} us / node")
