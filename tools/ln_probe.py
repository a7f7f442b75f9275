"""LayerNorm + SiLU over ~10^6 narrow rows (the conv layers' per-pixel channel LayerNorm, cnn_atari.py:78-86): us per launch and
GB/s of the forward (8 B / element) and backward (16 B / element) op at the conv channel counts. DV3_NO_LN_NARROW=1 selects the
warp-per-row kernels for comparison.   python tools/ln_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dreamer_v3_b200 import ops  # noqa: E402

tag = "warp-per-row" if os.environ.get("DV3_NO_LN_NARROW") else "default"
for rows, D in ((1 << 20, 32), (1 << 20, 48), (1 << 20, 64), (1 << 20, 96), (1 << 18, 192), (1 << 18, 128), (1 << 18, 256), (1 << 16, 384),
                (16384, 256), (16384, 640), (16384, 1024)):
    x = torch.randn(rows, D, device="cuda")
    dy = torch.randn(rows, D, device="cuda")
    g = torch.ones(D, device="cuda")
    b = torch.zeros(D, device="cuda")
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    bwd = ops.ln_silu_bwd(dy, x, g, b, dg, db)
    for name, fn, bytes_per in (("fwd", lambda: ops.ln_silu(x, g, b), 8), ("bwd", bwd, 12)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        n = 10
        for _ in range(n):
            fn()
        t.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(t) / n * 1e3
        print(f"{tag} rows={rows:8d} D={D:4d} {name}: {us:8.1f} us  {rows * D * bytes_per / us / 1e3:7.0f} GB/s")
