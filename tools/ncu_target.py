"""Small profiling target: one world-model update (observe forward + losses + BPTT) of a bench workload in the tensor-core mode,
non-graph path, so that `ncu -k regex:...` sees the individual kernels.  python tools/ncu_target.py c3 [passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_cfg, synthetic_sample  # noqa: E402
from dreamer_v3_b200.engine import Engine  # noqa: E402
from dreamer_v3_b200.spec import init_params  # noqa: E402

wid = sys.argv[1] if len(sys.argv) > 1 else "c3"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = make_cfg(wid)
e = Engine(cfg, compute_mode=1)
e.load_params(init_params(cfg, seed=0))
e.set_sample(synthetic_sample(cfg, 0))
e.fill_noise(1, 0, 3)
for _ in range(passes):
    e.observe_forward()
    e.wm_loss_backward()
torch.cuda.synchronize()
print("ok", wid, e.kernel_launch_count(), "launches")
e.close()
