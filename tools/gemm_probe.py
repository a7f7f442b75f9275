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
run = (lambda: ops.gemm(A, B, bool(ta), bool(tb), C_=C, mode=mode)) if mode != 3 else ops.gemm_weight_shadow(A, B, bool(ta), C_=C)
for _ in range(3):
    run()
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    run()
t.record()
torch.cuda.synchronize()
ms = s.elapsed_time(t) / 10
print(f"gemm {M}x{N}x{K} mode {mode} ta={ta} tb={tb}: {ms * 1e3:.1f} us, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
