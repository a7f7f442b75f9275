"""Generates tests/golden/nano_vec_disc.npz with the ORACLE (the reference itself cannot run in this image:
TensorFlow is not installable). It is a regression pin of the oracle and a fixed target for
the CUDA engine, not a reference-generated vector.   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dreamer_v3_b200.spec import EngineConfig, init_params  # noqa: E402
from oracle.dreamer_oracle import Oracle, make_synthetic_sample  # noqa: E402

cfg = EngineConfig.make("nano", A=2, discrete=True, B=3, T=5, H=4)
P = init_params(cfg, seed=11, mode="random")
sample = make_synthetic_sample(cfg, seed=12)
o = Oracle(cfg, P, dtype=torch.float64)
res = o.train_world_model_one_step(sample, rng=np.random.default_rng(13), margin=1e-2, apply=False)
fwd = res["WORLD_MODEL_forward_train_outs"]
out = {"noise/u_post": o.last_noise["u_post"].astype(np.float32)}
out.update({"param/" + k: v for k, v in P.items()})
out.update({"sample/" + k: v for k, v in sample.items()})
out["out/z_indices"] = fwd["z_indices_BxT"].astype(np.int32)
out["out/reward_two_hot_k"] = res["reward_two_hot_k"].astype(np.int32)
out["out/L_total"] = np.float64(res["WORLD_MODEL_L_total"].item())
out["out/h_states"] = fwd["h_states_BxT"].numpy().astype(np.float32)
out.update({"grad/" + k: v.numpy().astype(np.float32) for k, v in res["grads"].items()})
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "nano_vec_disc.npz"), **out)
print("wrote", len(out), "arrays")
