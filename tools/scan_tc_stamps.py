"""Per-stage timing of the tensor-core observe scans (dv3_scan_tc.cu): globaltimer stamps of CTA 0 after the operand staging,
after the last accumulator of the stage has been drained, and after the grid barrier that closes the stage.
    python tools/scan_tc_stamps.py c4"""
import os
import sys

os.environ["DV3_SCAN_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import make_cfg, synthetic_sample  # noqa: E402
from dreamer_v3_b200.engine import Engine  # noqa: E402
from dreamer_v3_b200.spec import init_params  # noqa: E402

wid = sys.argv[1] if len(sys.argv) > 1 else "c4"
cfg = make_cfg(wid)
e = Engine(cfg, compute_mode=1)
e.load_params(init_params(cfg, seed=0))
e.set_sample(synthetic_sample(cfg, 0))
e.fill_noise(1, 0, 3)
T = cfg.T


def stamps():
    st = e.tensor("debug.scan_stamps").cpu().numpy().astype(np.int64)
    return ((st[0::2] & 0xFFFFFFFF) | (st[1::2] << 32)).astype(np.float64)


for _ in range(3):
    e.observe_forward()
torch.cuda.synchronize()
st = stamps()
# forward: 1 + T * (3 stages x 3 stamps + 2 for the sampling stage)
d = np.diff(st[: 1 + T * 11]) / 1e3
d = d.reshape(T, 11)[2:].mean(0)
names = ["S1 stage", "S1 mma+drain", "S1 barrier", "S2 stage", "S2 mma+drain", "S2 barrier", "S3 stage", "S3 mma+drain", "S3 barrier",
         "S4 sample+gather", "S4 barrier"]
print(wid, "forward us/step:", {n: round(float(v), 2) for n, v in zip(names, d)}, "total", round(float(d.sum()), 2))


def prof(tag):
    """per-CTA cycle totals of the worker warps (thread 0) over the whole kernel, as us per step at 1.9 GHz"""
    st = e.tensor("debug.scan_stamps").cpu().numpy().astype(np.int64)
    v = ((st[0::2] & 0xFFFFFFFF) | (st[1::2] << 32)).astype(np.float64)[2048:2048 + 8 * 148].reshape(148, 8) / 1.9e3 / T
    cats = ["-", "rows", "batch", "publish", "acc wait", "reds", "barrier", "extra"]
    tot = v.sum(1)
    print(wid, tag, "mean over CTAs us/step:", {c: round(float(x), 2) for c, x in zip(cats, v.mean(0))}, "total", round(float(tot.mean()), 2))
    for name, col in (("rows", 1), ("batch", 2), ("publish", 3), ("acc wait", 4), ("barrier", 6)):
        i = int(v[:, col].argmax())
        print("   max", name, "at CTA", i, {c: round(float(x), 2) for c, x in zip(cats, v[i])})


prof("forward")
e.wm_loss_backward()
torch.cuda.synchronize()
st = stamps()
d = np.diff(st[: 1 + T * 14]) / 1e3
d = d.reshape(T, 14)[2:-1].mean(0)
names = ([f"B{i} {w}" for i in (1, 2) for w in ("stage", "mma+drain", "barrier")] + ["B3a gru-bwd", "B3a barrier"]
         + [f"B{i} {w}" for i in (3, 4) for w in ("stage", "mma+drain", "barrier")])
print(wid, "backward us/step:", {n: round(float(v), 2) for n, v in zip(names, d)}, "total", round(float(d.sum()), 2))
prof("backward")
e.close()
