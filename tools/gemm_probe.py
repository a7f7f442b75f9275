"""Runs one GEMM shape through libdv3 (for ncu captures): python tools/gemm_probe.py M N K mode [ta tb]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dreamer_v3_b200 import ops  # noqa: E402

M, N, K, mode = (int(x) for x in sys.argv[1:5])
ta, tb = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (0, 0)
A = torch.randn((K, M) if ta else (M, K), device="cuda")
B = torch.randn((N, K) if tb else (K, N), device="cuda")
C = torch.empty(M, N, device="cuda")
if mode == 5:   # weight-gradient contraction X^T dY on bf16 operands as stored: A [K, M], B [K, N]
    A16, B16 = ops.to_bf16(torch.randn(K, M, device="cuda")), ops.to_bf16(torch.randn(K, N, device="cuda"))
    run = lambda: ops.gemm_bf16_wgrad(A16, B16, C_=C)  # noqa: E731
elif mode == 4:   # TMA-fed kernel: A in bf16 (as the producer kernels write it), B as bf16 shadow
    run = ops.gemm_bf16(ops.to_bf16(A), B, C_=C)
elif mode == 3:
    run = ops.gemm_weight_shadow(A, B, bool(ta), C_=C)
else:
    run = lambda: ops.gemm(A, B, bool(ta), bool(tb), C_=C, mode=mode)  # noqa: E731
for _ in range(3):
    run()
torch.cuda.synchronize()
# back-to-back replays of a captured graph: launch overhead of the Python / ctypes call is not in the number
REPS = 20
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    run()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for _ in range(REPS):
            run()
    gr.replay()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    gr.replay()
    t.record()
    torch.cuda.synchronize()
ms = s.elapsed_time(t) / REPS
print(f"gemm {M}x{N}x{K} mode {mode} ta={ta} tb={tb}: {ms * 1e3:.1f} us, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
