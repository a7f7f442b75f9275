"""Runs the oracle-vs-engine comparison for a few configs on the GPU and prints every deviation
(debug aid; the pass/fail version is tests/test_gpu_engine.py)."""
import argparse
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import torch  # noqa: E402

from _parity import ParityRun  # noqa: E402
from dreamer_v3_b200.spec import EngineConfig  # noqa: E402

CASES = {
    "nano-vec-disc": dict(size="nano", A=2, discrete=True, B=3, T=5, H=4),
    "nano-vec-box": dict(size="nano", A=3, discrete=False, B=3, T=5, H=4),
    "nano-img-disc": dict(size="nano", A=4, discrete=True, image_obs=True, B=2, T=3, H=3),
    "micro-img-box": dict(size="micro", A=2, discrete=False, image_obs=True, B=2, T=4, H=3),
    "xxs-vec-disc": dict(size="XXS", A=2, discrete=True, B=4, T=8, H=5),
    "xxs-b8": dict(size="XXS", A=3, discrete=True, B=8, T=16, H=3),
    "xs-cartpole": dict(size="XS", A=2, discrete=True, B=16, T=64, H=15, obs_dim=4),
    "xs-pendulum": dict(size="XS", A=1, discrete=False, B=16, T=64, H=15, obs_dim=3),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="nano-vec-disc,nano-vec-box,nano-img-disc")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--top", type=int, default=400)
    ap.add_argument("--thresh", type=float, default=1e-4)
    ap.add_argument("--margin", type=float, default=1e-3)
    args = ap.parse_args()
    for name in args.cases.split(","):
        kw = dict(CASES[name])
        size = kw.pop("size")
        cfg = EngineConfig.make(size, **kw)
        print(f"===== {name}  {cfg.asdict()}", flush=True)
        t0 = time.time()
        try:
            run = ParityRun(cfg, steps=args.steps, compute_mode=args.mode, margin=args.margin)
        except Exception:
            traceback.print_exc()
            continue
        print(f"  time {time.time() - t0:.1f}s  exact mismatches: {run.exact}")
        bad = [(k, v) for k, v in run.worst(args.top) if not (v <= args.thresh)]
        print(f"  {len(run.report)} tensors compared, {len(bad)} above {args.thresh}")
        for k, v in run.worst(args.top):
            flag = "  <-- " if not (v <= args.thresh) else ""
            if flag or args.top < 50:
                print(f"    {k:60s} {v:.3e}{flag}")
        print("  worst 8:", [(k, f"{v:.2e}") for k, v in run.worst(8)])
        run.close()
        torch.cuda.synchronize()
