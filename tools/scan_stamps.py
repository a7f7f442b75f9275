"""Per-stage timing of the persistent observe-scan forward kernel (CTA 0 globaltimer stamps)."""
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
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # 0 fp32, 1 bf16 tensor-core mode
e = Engine(cfg, compute_mode=mode)
e.load_params(init_params(cfg, seed=0))
e.set_sample(synthetic_sample(cfg, 0))
e.fill_noise(1, 0, 3)
for _ in range(3):
    e.observe_forward()
torch.cuda.synchronize()
st = e.tensor("debug.scan_stamps").cpu().numpy().astype(np.int64)
st = (st[0::2] & 0xFFFFFFFF) | (st[1::2] << 32)
T = cfg.T
d = np.diff(st[: T * 8 + 1].astype(np.float64)) / 1e3  # us
d = d.reshape(T, 8)
names = ["A work", "A sync", "B work", "B sync", "C work", "C sync", "D work", "D sync"]
ex = None

print("per-step mean us:", {n: round(float(d[2:, i].mean()), 2) for i, n in enumerate(names)}, "step total", round(float(d[2:].sum(1).mean()), 2))


# backward scan (same stamp buffer, written by the next kernel)
e.wm_loss_backward()
torch.cuda.synchronize()
st = e.tensor("debug.scan_stamps").cpu().numpy().astype(np.int64)
st = (st[0::2] & 0xFFFFFFFF) | (st[1::2] << 32)
d = np.diff(st[: T * 8 + 1].astype(np.float64)) / 1e3
d = d.reshape(T, 8)
if mode == 1 and not os.environ.get("DV3_NO_CLUSTER_BWD"):
    # cluster kernel: stamps close S1, S2, S3, S4 and follow each barrier; S5 of step t is counted with S1 of step t-1
    names = ["S5+S1 work", "sync1", "S2 work", "sync2", "S3 work", "sync3", "S4 work", "sync4"]
    print("backward (cluster) per-step mean us:", {n: round(float(d[2:-1, i].mean()), 2) for i, n in enumerate(names)}, "step total", round(float(d[2:-1].sum(1).mean()), 2))
    # sub-stamps inside "S5+S1": after the LayerNorm backward of S5, after dz_in, after the categorical backward of S1, after its broadcast
    sub = st[T * 8 + 16: T * 8 + 16 + T * 4].astype(np.float64).reshape(T, 4)
    s8 = st[: T * 8 + 1].astype(np.float64)
    rows = []
    for k in range(2, T - 1):   # step index k (t = T-1-k): S5 of step k starts at stamp 8k+8, S1 of step k+1 ends at stamp 8(k+1)+1
        t0 = s8[8 * k + 8]
        rows.append([(sub[k, 0] - t0), (sub[k, 1] - sub[k, 0]), (sub[k + 1, 2] - sub[k, 1]), (sub[k + 1, 3] - sub[k + 1, 2]), (s8[8 * (k + 1) + 1] - sub[k + 1, 3])])
    m = np.array(rows).mean(0) / 1e3
    print("  S5 LN bwd %.2f, S5 dz_in %.2f, S1 categorical bwd %.2f, S1 broadcast %.2f, step loads %.2f us" % tuple(m))
elif os.environ.get("DV3_SCAN_BWD_V1"):
    names = ["S1 work", "S1 sync", "S2 work", "S2 sync", "S3 work", "S3 sync", "S4 work", "S4 sync"]
    print("backward (four-stage) per-step mean us:", {n: round(float(d[2:-1, i].mean()), 2) for i, n in enumerate(names)}, "step total", round(float(d[2:-1].sum(1).mean()), 2))
else:   # two-stage kernel: stamps 1-3 close X(t), 4 follows its barrier, 5-7 close Y(t), 8 follows its barrier
    m = d[2:-1].mean(0)
    print("backward (two-stage) per-step mean us:", {"X work": round(float(m[0] + m[1] + m[2]), 2), "X sync": round(float(m[3]), 2),
                                                     "Y work": round(float(m[4] + m[5] + m[6]), 2), "Y sync": round(float(m[7]), 2)},
          "step total", round(float(d[2:-1].sum(1).mean()), 2))
