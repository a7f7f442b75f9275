"""Curiosity (Plan2Explore disagree nets) at the benchmark shape: C1 (CartPole / XS / Discrete, B16 x T64 x H15) with the
reference's 8 nets, tensor-core mode. Default: updates/s of the graph-replayed update with and without curiosity (CUDA events).
`target`: one non-graph actor / critic step so that `ncu -k regex:disagree` sees the kernels of csrc/dv3_disagree.cu.
  python tools/curiosity_prof.py            |  ncu ... python tools/curiosity_prof.py target"""
import dataclasses
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_cfg, synthetic_sample  # noqa: E402
from dreamer_v3_b200.engine import Engine  # noqa: E402
from dreamer_v3_b200.spec import init_params  # noqa: E402

base = make_cfg("c1")


def build(nets):
    cfg = dataclasses.replace(base, num_curiosity_nets=nets, intrinsic_rewards_scale=0.1)
    e = Engine(cfg, compute_mode=1)
    e.load_params(init_params(cfg, seed=0))
    e.set_sample(synthetic_sample(cfg, 0))
    return cfg, e


if len(sys.argv) > 1 and sys.argv[1] == "target":
    cfg, e = build(8)
    e.fill_noise(1, 0, 3)
    for _ in range(2):
        e.observe_forward()
        e.dream_forward()
        e.ac_loss_backward(3)
        e.disagree_loss_backward()
    torch.cuda.synchronize()
    R, Z = (cfg.H + 1) * cfg.N, cfg.Z
    print("ok rows", R, "std kernel bytes", 8 * R * Z * 4, "loss kernel bytes per net", 3 * cfg.H * cfg.N * Z * 4)
    e.close()
    sys.exit(0)

for nets in (0, 8):
    cfg, e = build(nets)
    for _ in range(5):
        e.train_update(noise_seed=7, use_graph=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 30
    for _ in range(n):
        e.train_update(noise_seed=7, use_graph=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"nets {nets}: {ms:.3f} ms per update = {1e3 / ms:.1f} updates/s", "| DISAGREE_L_total", float(e.tensor("disagree.scalars")[0]) if nets else "-")
    e.close()
