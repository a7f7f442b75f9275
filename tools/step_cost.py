"""In-graph cost of one dreamed step: times the graph-replayed actor-critic update at two horizons."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_cfg, synthetic_sample  # noqa: E402
from dreamer_v3_b200.engine import Engine  # noqa: E402
from dreamer_v3_b200.spec import init_params  # noqa: E402

wid = sys.argv[1] if len(sys.argv) > 1 else "c2"
res = {}
for H in (15, 5):
    cfg = make_cfg(wid, H=H)
    e = Engine(cfg, compute_mode=1)
    e.load_params(init_params(cfg, seed=0))
    e.set_sample(synthetic_sample(cfg, 0))
    for _ in range(5):
        e.train_world_model_step(1)
        e.train_actor_critic_step(1)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    wm = ac = 0.0
    for _ in range(20):
        ev[0].record(); e.train_world_model_step(1); ev[1].record(); e.train_actor_critic_step(1); ev[2].record()
        torch.cuda.synchronize()
        wm += ev[0].elapsed_time(ev[1]); ac += ev[1].elapsed_time(ev[2])
    res[H] = (wm / 20, ac / 20)
    print(f"H={H}: world-model step {wm / 20:.3f} ms, actor-critic step {ac / 20:.3f} ms")
    e.close()
print(f"actor-critic cost per dreamed step (fwd + bwd incl. its share of the head rows): {(res[15][1] - res[5][1]) / 10 * 1e3:.1f} us")
