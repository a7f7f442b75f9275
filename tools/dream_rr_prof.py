"""Per-phase cycle totals of the row-resident dream kernel (dv3_dream_rr.cu), first consumer warp of every CTA:
    python tools/dream_rr_prof.py c2"""
import os
import sys

os.environ["DV3_SCAN_STAMPS"] = "1"
os.environ["DV3_DREAM_RR"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import make_cfg, synthetic_sample  # noqa: E402
from dreamer_v3_b200.engine import Engine  # noqa: E402
from dreamer_v3_b200.spec import init_params  # noqa: E402

wid = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = make_cfg(wid, B=int(sys.argv[2]) if len(sys.argv) > 2 else 16)   # fewer rows = fewer CTAs streaming the same weights
e = Engine(cfg, compute_mode=1)
e.load_params(init_params(cfg, seed=0))
e.set_sample(synthetic_sample(cfg, 0))
e.fill_noise(1, 0, 3)
e.observe_forward()
for _ in range(3):
    e.dream_forward(start_from_observe=True)
torch.cuda.synchronize()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record(); e.dream_forward(start_from_observe=True); s1.record(); torch.cuda.synchronize()
print(wid, "B", cfg.B, "CTAs", cfg.N // 16, "dream_forward ms", round(s0.elapsed_time(s1), 3))
ctas = cfg.N // 16
st = e.tensor("debug.scan_stamps").cpu().numpy().astype(np.int64)
v = ((st[0::2] & 0xFFFFFFFF) | (st[1::2] << 32)).astype(np.float64)[2048:2048 + 16 * ctas].reshape(ctas, 16)[:, :12] / 1.9e3 / cfg.H
names = ["gathers", "barrier waits", "actor tiles", "actor epilogue", "action + pre-GRU layer", "GRU tiles", "GRU cell", "dyn tiles",
         "q epilogue", "logit tiles", "sampling epilogue", "-"]
print("us per step, mean over CTAs:", {n: round(float(x), 2) for n, x in zip(names, v.mean(0)) if n != "-"}, "total", round(float(v.sum(1).mean()), 2))
print("             max over CTAs:", {n: round(float(x), 2) for n, x in zip(names, v.max(0)) if n != "-"})
e.close()
