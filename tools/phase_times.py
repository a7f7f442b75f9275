"""Device time per phase of one update on the non-graph path (DV3_PHASE_TIMES instrumentation in the engine)."""
import os
import sys

os.environ["DV3_PHASE_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_cfg, synthetic_sample  # noqa: E402
from dreamer_v3_b200.engine import Engine  # noqa: E402
from dreamer_v3_b200.spec import init_params  # noqa: E402

wid = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = make_cfg(wid)
e = Engine(cfg, compute_mode=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
e.load_params(init_params(cfg, seed=0))
e.set_sample(synthetic_sample(cfg, 0))
for it in range(3):
    e.fill_noise(1, it, 3)
    e.observe_forward()
    e.wm_loss_backward()
    e.dream_forward()
    e.ac_loss_backward()
    torch.cuda.synchronize()
    if it < 2:
        e.sync()   # prints and drops the warm-up marks
        print("---- (warm-up above) ----", file=sys.stderr)
e.sync()
