"""World-model gradient of one update at a bench workload in tensor-core mode, saved to a file (run once per variant of the
backward scan, e.g. with and without DV3_NO_CLUSTER_BWD=1) or compared: python tools/bwd_cluster_check.py c2 out.npy |
python tools/bwd_cluster_check.py --cmp a.npy b.npy"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    d = np.abs(a - b)
    print("flat grad: n", a.size, "max|a|", float(np.abs(a).max()), "max|a-b|", float(d.max()), "rel", float(d.max() / np.abs(a).max()),
          "rel l2", float(np.linalg.norm(a - b) / np.linalg.norm(a)))
    n = 16
    for i, (x, y) in enumerate(zip(np.array_split(a, n), np.array_split(b, n))):
        print(f"  chunk {i:2d}: max|a| {np.abs(x).max():.3e} rel {np.abs(x - y).max() / max(np.abs(x).max(), 1e-30):.3e}")
    sys.exit(0)

import torch
from bench import make_cfg, synthetic_sample
from dreamer_v3_b200.engine import Engine
from dreamer_v3_b200.spec import init_params

cfg = make_cfg(sys.argv[1])
e = Engine(cfg, compute_mode=1)
e.load_params(init_params(cfg, seed=0))
e.set_sample(synthetic_sample(cfg, 0))
e.fill_noise(1, 0, 3)
e.observe_forward()
e.wm_loss_backward()
torch.cuda.synchronize()
np.save(sys.argv[2], e.group_buffer("world_model", "grad").detach().cpu().numpy())
print("saved", sys.argv[2])
