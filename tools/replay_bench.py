#!/usr/bin/env python
"""Device replay sampler measured on its HBM roofline, next to the reference algorithm on the host.

One B16 x T64 sample of 64x64x3 frames (the Atari / DMC shape of BASELINE configs 3-5): the gather kernel
`replay_gather_kernel` reads 12 288 B (uint8 store) or 49 152 B (float32 store) and writes 49 152 B per timestep.
Timed with CUDA events over graph-free back-to-back launches on rotating output buffers (> L2); the host plan
(`EpisodeReplayBuffer.plan`) is timed separately.  CPU figure: the oracle restatement of the reference's list-packing
`sample()` (kind "port"; the reference's own buffer is pure NumPy but /root/reference is not on the GPU box).

    python tools/replay_bench.py            # one JSON line
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def run(B=16, T=64, iters=50, cpu=True):
    from dreamer_v3_b200.utils.episode import Episode
    from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
    from oracle.replay_oracle import ReplayOracle
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    rng = np.random.default_rng(0)
    res = {"workload": f"B{B} x T{T} rows of 64x64x3 frames", "peak_gbs": hbm,
           "peak_source": "measured" if os.path.exists(peaks) else "fallback"}
    episodes = []
    for e in range(24):
        n = int(rng.integers(100, 600))
        episodes.append((f"e{e}", rng.integers(0, 256, size=(n + 1, 64, 64, 3), dtype=np.uint8),
                         np.eye(6, dtype=np.float32)[rng.integers(0, 6, size=n)],
                         rng.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32), bool(rng.random() < 0.7)))
    for store in ("u8", "f32"):
        buf = EpisodeReplayBuffer(capacity=20000, normalize_uint8=True)
        for id_, obs, act, rew, term in episodes:
            o = obs if store == "u8" else (obs.astype(np.float32) / 128.0 - 1.0)
            buf.add(Episode(id_, observations=list(o), actions=list(act), rewards=list(rew), is_terminated=term))
        outs = [{"obs": torch.empty((B, T, 64, 64, 3), device="cuda")} for _ in range(4)]  # 4 x 50 MB > L2
        np.random.seed(1)
        t0 = time.perf_counter()
        for _ in range(20):
            buf.plan(B, T)
        plan_ms = (time.perf_counter() - t0) / 20 * 1e3
        for i in range(5):
            buf.sample(B, T, out=outs[i % 4])
        # kernel alone: replay the gather on prebuilt plans
        import ctypes as C
        plans = [torch.from_numpy(buf.plan(B, T)).cuda() for _ in range(4)]
        full = [buf.sample(B, T, out=outs[i]) for i in range(4)]
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(iters):
            o = full[i % 4]
            buf.lib.dv3_replay_gather(p(buf._obs_store), int(buf._obs_u8), 1, buf._obs_elems, p(buf._act_store), buf._A,
                                      p(buf._rew_store), p(plans[i % 4]), B * T, p(o["obs"]), p(o["actions"]), p(o["rewards"]),
                                      p(o["is_first"]), p(o["is_last"]), p(o["is_terminated"]), st)
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) / iters * 1e3
        # a plain device copy of as many bytes as the sample holds (what a STREAM copy reaches at this size)
        src = torch.empty((B * T * 12288,), dtype=torch.float32, device="cuda")
        dsts = [o["obs"].view(-1) for o in outs]
        for i in range(4):
            dsts[i].copy_(src)
        torch.cuda.synchronize()
        ev0.record()
        for i in range(iters):
            dsts[i % 4].copy_(src)
        ev1.record()
        torch.cuda.synchronize()
        copy_us = ev0.elapsed_time(ev1) / iters * 1e3
        per_ts = 12288 * (1 if store == "u8" else 4) + 12288 * 4 + 6 * 4 * 2 + 4 * 4 + 16
        gbs = B * T * per_ts / (us * 1e-6) / 1e9
        # end to end: host plan + H2D of the plan + gather, synchronised
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(20):
            buf.sample(B, T, out=outs[i % 4])
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / 20 * 1e3
        res[store] = {"gather_us": round(us, 2), "algorithmic_bytes": B * T * per_ts, "achieved_gbs": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / hbm, 3),
                      "torch_copy_same_output_us": round(copy_us, 2), "host_plan_ms": round(plan_ms, 3),
                      "sample_call_ms": round(e2e_ms, 3), "samples_per_s": round(1e3 / e2e_ms, 1)}
    if cpu:
        o = ReplayOracle(20000)
        for id_, obs, act, rew, term in episodes:
            o.add(id_, list(obs.astype(np.float32) / 128.0 - 1.0), list(act), list(rew), term)
        np.random.seed(1)
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            o.sample(B, T)
        ms = (time.perf_counter() - t0) / n * 1e3
        res["cpu_baseline"] = {"sample_call_ms": round(ms, 2), "samples_per_s": round(1e3 / ms, 2), "cores": 1, "kind": "port",
                               "sample": f"{n} sample() calls of the restated list-packing loop + np.array(), float32 frames"}
    return res


if __name__ == "__main__":
    print(json.dumps(run()))
