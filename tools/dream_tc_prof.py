"""Per-CTA cycle totals of the worker warps of the persistent dream kernel (dv3_dream_tc.cu):  python tools/dream_tc_prof.py c2"""
import os
import sys

os.environ["DV3_SCAN_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import make_cfg, synthetic_sample  # noqa: E402
from dreamer_v3_b200.engine import Engine  # noqa: E402
from dreamer_v3_b200.spec import init_params  # noqa: E402

wid = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = make_cfg(wid)
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
print(wid, "dream_forward ms", round(s0.elapsed_time(s1), 3))
st = e.tensor("debug.scan_stamps").cpu().numpy().astype(np.int64)
v = ((st[0::2] & 0xFFFFFFFF) | (st[1::2] << 32)).astype(np.float64)[2048:2048 + 16 * 148].reshape(148, 16) / 1.9e3 / cfg.H
cats = ["-", "elementwise", "-", "acc wait", "epilogue", "barrier", "stage fence", "-"]
print("mean over CTAs us/step:", {c: round(float(x), 2) for c, x in zip(cats, v[:, :8].mean(0)) if c != "-"}, "total", round(float(v[:, :8].sum(1).mean()), 2))
w = v[:, 8:]
print("per stage of the step, work until the barrier arrive, us: max over CTAs", [round(float(x), 2) for x in w.max(0)], "sum", round(float(w.max(0).sum()), 2))
print("                                                        mean over CTAs", [round(float(x), 2) for x in w.mean(0)])
i = int(w.sum(1).argmax())
print("   busiest CTA", i, [round(float(x), 2) for x in w[i]], {c: round(float(x), 2) for c, x in zip(cats, v[i, :8]) if c != "-"})
e.close()
