#!/usr/bin/env python
"""Benchmark of the DreamerV3 train update (train_world_model_one_step + train_actor_and_critic_one_step,
B16 x T64, H15) on B200, contract in the task statement:

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0). `value` = device-timed whole-job throughput with inputs resident in HBM;
`e2e` = the same update through the reference-facing Python API with pinned HOST inputs (H2D + D2H inside the
timed region). Weak scaling: every rank trains on its own B=16 replay rows, gradients all-reduced over NCCL;
value counts B16xT64 batches consumed per second over all ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL prints its version banner on stdout at the VERSION and WARN debug levels, so
# the banner level is switched off and whatever NCCL still logs goes to stderr
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train_updates_per_sec_B16xT64_H15"

# BASELINE.json configs: id -> EngineConfig.make kwargs
WORKLOADS = {
    "c1": dict(name="CartPole-v1 / XS / Discrete(2) / vector obs", size="XS", A=2, discrete=True, obs_dim=4),
    "c2": dict(name="Pendulum-v1 / XS / Box(1) / vector obs", size="XS", A=1, discrete=False, obs_dim=3),
    "c3": dict(name="Atari100k Pong / S / Discrete(6) / 64x64x3", size="S", A=6, discrete=True, image_obs=True),
    "c4": dict(name="DMC vision / M / Box(6) / 64x64x3", size="M", A=6, discrete=False, image_obs=True),
    "L": dict(name="size L / Discrete(18) / 64x64x3", size="L", A=18, discrete=True, image_obs=True),
    "c5": dict(name="size XL / Discrete(18) / 64x64x3", size="XL", A=18, discrete=True, image_obs=True),
}
# algorithmic GFLOP per update
GFLOP_PER_UPDATE = {"c1": 150.0, "c2": 235.0, "c3": 827.0, "c4": 2319.0, "L": 3506.0, "c5": 8756.0}


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the roofline kernel from the committed ncu --set full
# capture (profiles/), keyed by (M, K, N, mode); None when no capture exists for the shape
TRAFFIC_BYTES = {(16384, 1280, 256, 1): 42616320 + 142080,   # profiles/r01_gemm_tma_ncu.txt (the fp32 C tile stays in L2 past kernel end)
                 (16384, 1280, 256, 0): None}


def make_cfg(wid, B=16, T=64, H=15):
    from dreamer_v3_b200.spec import EngineConfig
    kw = {k: v for k, v in WORKLOADS[wid].items() if k not in ("name", "size")}
    return EngineConfig.make(WORKLOADS[wid]["size"], B=B, T=T, H=H, **kw)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_sample(cfg, seed):
    from oracle.dreamer_oracle import make_synthetic_sample  # data generator only (no oracle compute)
    return make_synthetic_sample(cfg, seed=seed)


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference algorithm on the box's host cores: the oracle restatement (TensorFlow is not installable in
    this image), PyTorch-CPU fp32 eager, same 64-step / 15-step Python loops + autograd, all host
    threads. Each step is one full update on a bounded sample of the workload (B rows reduced when one update
    would take minutes) and scaled to B=16."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dreamer_v3_b200.spec import init_params
    from oracle.dreamer_oracle import Oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wid = args.workload
    rows = {"c1": 16, "c2": 16, "c3": 2, "c4": 1, "L": 1, "c5": 1}[wid] if args.ref_rows is None else args.ref_rows
    cfg = make_cfg(wid, B=rows)
    sample = synthetic_sample(cfg, 0)
    rng = np.random.default_rng(0)
    o = Oracle(cfg, init_params(cfg, seed=0), dtype=torch.float32)

    def step():
        res = o.train_world_model_one_step(sample, rng=rng)
        f = res["WORLD_MODEL_forward_train_outs"]
        o.train_actor_and_critic_one_step(f["h_states_BxT"], f["z_states_BxT"], sample["is_terminated"], H=cfg.H, rng=rng)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    dt16 = dt * (16.0 / rows)  # cost is linear in the number of replay rows
    val = 1.0 / dt16
    samp = f"{args.steps} full updates at B={rows} of 16 rows (x{16 // rows} scaled), T=64, H=15, all {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt16 * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[wid]["name"], "id": wid, "B_per_gpu": 16, "T": 64, "H": 15},
        "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": "port", "sample": samp},
        "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm, PyTorch-CPU restatement (TensorFlow unavailable in this image)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from dreamer_v3_b200.models.components import CNNAtari, ConvTransposeAtari, MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from dreamer_v3_b200.training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wid = args.workload
    w = WORKLOADS[wid]
    cfg = make_cfg(wid)
    B, T, H = cfg.B, cfg.T, cfg.H
    aspace = Discrete(w["A"]) if w["discrete"] else Box(-1.0, 1.0, (w["A"],))
    size = w["size"]
    if cfg.image_obs:
        enc, dec = CNNAtari(model_dimension=size), ConvTransposeAtari(model_dimension=size, gray_scaled=False)
    else:
        enc, dec = MLP(model_dimension=size), VectorDecoder(model_dimension=size, observation_space=Box(-1, 1, (cfg.obs_dim,)))
    wm = WorldModel(model_dimension=size, action_space=aspace, batch_length_T=T, encoder=enc, decoder=dec,
                    symlog_obs=not cfg.image_obs, compute_mode=args.mode)
    dm = DreamerModel(model_dimension=size, action_space=aspace, batch_size_B=B, batch_length_T=T, horizon_H=H,
                      world_model=wm, world_size=world, rank=rank)
    e = dm.engine
    dm.critic.init_ema()
    sample_np = synthetic_sample(cfg, seed=rank)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in sample_np.items()}
    h2d_bytes = sum(v.numel() * 4 for v in pinned.values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    host_scalars = torch.empty(24, dtype=torch.float32).pin_memory()
    d2h_bytes = host_scalars.numel() * 4
    hp = dict(horizon_H=H, gamma=0.997, lambda_=0.95, actor_grad_clip=100.0, critic_grad_clip=100.0,
              disagree_grad_clip=100.0, entropy_scale=3e-4, return_normalization_decay=0.99)

    def api_step(host_inputs):
        sample = pinned if host_inputs else None
        if sample is None:  # inputs already resident in the engine's in.* buffers
            sample = {}
        r = train_world_model_one_step(sample=sample, batch_size_B=B, batch_length_T=T, grad_clip=1000.0, world_model=wm)
        a = train_actor_and_critic_one_step(world_model_forward_train_outs=r["WORLD_MODEL_forward_train_outs"],
                                            is_terminated=e.tensor("in.is_terminated").reshape(-1), dreamer_model=dm, **hp)
        if host_inputs:
            host_scalars[:8].copy_(e.tensor("wm.scalars"), non_blocking=True)
            host_scalars[8:24].copy_(e.tensor("ac.scalars"), non_blocking=True)
        return r, a

    use_graph = world == 1 and not args.no_graph

    def device_step():
        if use_graph:
            e.train_update(noise_seed=1234, use_graph=True)
        else:
            api_step(False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        for _ in range(steps):
            flush.zero_()  # L2 flush (256 MiB > 126 MB L2) between timed iterations, outside the timed events
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); t.record()
            evs.append((s, t))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / steps

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg ----
    e.set_sample(pinned)
    if args.breakdown:
        e.fill_noise(1, 0, 3)
        phases = [("observe_forward", e.observe_forward), ("wm_loss_backward", e.wm_loss_backward),
                  ("wm_optimizer", lambda: e.optimizer_step("world_model", 1000.0, 1e-4, 1e-8)),
                  ("dream_forward", lambda: e.dream_forward(start_from_observe=True)),
                  ("ac_loss_backward", lambda: e.ac_loss_backward(3)),
                  ("ac_optimizers", lambda: (e.optimizer_step("actor", 100.0, 3e-5, 1e-5), e.optimizer_step("critic", 100.0, 3e-5, 1e-5, 0.98)))]
        out = {}
        for rep in range(4):
            for name, fn in phases:
                torch.cuda.synchronize()
                l0 = e.kernel_launch_count()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(); fn(); s1.record(); torch.cuda.synchronize()
                out[name] = (round(s0.elapsed_time(s1), 3), e.kernel_launch_count() - l0)
        print(json.dumps({"breakdown_ms_launches": out, "total_ms": round(sum(v[0] for v in out.values()), 3)}), flush=True)
        return
    for _ in range(args.warmup if args.quick else max(args.warmup, 3)):
        device_step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = e.kernel_launch_count()
    wall0 = time.perf_counter()
    ms = timed(device_step, args.steps)
    barrier()
    wall = time.perf_counter() - wall0
    launches = e.kernel_launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    value = world * 1e3 / ms

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms, "value": value, "gpu_launches": int(launches)}), flush=True)
        return
    # ---- end-to-end leg (public API, pinned host inputs, D2H of the losses) ----
    for _ in range(3):
        api_step(True)
    barrier()

    def e2e_step():
        api_step(True)
        torch.cuda.current_stream().synchronize()  # the caller reads the losses
    ms_e2e = max_over_ranks(timed(e2e_step, args.steps))
    barrier()
    final = {"wm_total": float(host_scalars[6]), "critic_total": float(host_scalars[8]), "actor_total": float(host_scalars[11])}

    # ---- the reference's inner loop with the device replay sampler: buffer.sample() -> WM step -> AC step,
    # run_experiment.py:410-472; the sample is gathered from the HBM-resident episode store straight into in.* ----
    replay_leg = None
    if world == 1:
        from dreamer_v3_b200.utils.episode import Episode
        from dreamer_v3_b200.utils.episode_replay_buffer import EpisodeReplayBuffer
        rng = np.random.default_rng(3)
        buf = EpisodeReplayBuffer(capacity=20000)
        for i in range(40):
            n = int(rng.integers(50, 500))
            if cfg.image_obs:
                obs = rng.integers(0, 256, size=(n + 1, cfg.img_hw, cfg.img_hw, cfg.img_c), dtype=np.uint8)
            else:
                obs = rng.standard_normal((n + 1, cfg.obs_dim)).astype(np.float32)
            act = (np.eye(cfg.A, dtype=np.float32)[rng.integers(0, cfg.A, size=n)] if cfg.discrete
                   else rng.uniform(-1, 1, size=(n, cfg.A)).astype(np.float32))
            buf.add(Episode(f"e{i}", observations=list(obs), actions=list(act), rewards=list(rng.standard_normal(n).astype(np.float32)),
                            is_terminated=bool(rng.random() < 0.7)))
        outs = {"obs": e.tensor("in.obs"), "actions": e.tensor("in.actions"), "rewards": e.tensor("in.rewards"),
                "is_first": e.tensor("in.is_first"), "is_terminated": e.tensor("in.is_terminated")}
        np.random.seed(0)

        def replay_step():
            buf.sample(B, T, out=outs)
            api_step(False)
            host_scalars[:8].copy_(e.tensor("wm.scalars"), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3):
            replay_step()
        ms_rp = timed(replay_step, args.steps)
        replay_leg = {"value": 1e3 / ms_rp, "unit": "updates/s", "ms_per_step": ms_rp,
                      "what": "EpisodeReplayBuffer.sample (device store, dv3_replay_gather) + both update steps + D2H of the losses"}

    # ---- RSSM scan step latency (observe scan forward / T) ----
    def obs_fwd():
        e.observe_forward()
    for _ in range(3):
        obs_fwd()
    scan_us = max_over_ranks(timed(obs_fwd, max(5, args.steps // 2))) * 1e3 / T

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    gflop = GFLOP_PER_UPDATE[wid]
    line = {
        "metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == 0 else "bf16", "data": "synthetic",
        "precision_note": ("fp32 everywhere" if args.mode == 0 else "bf16 operands / fp32 accumulate on tcgen05 for contractions with >= 128 rows; fp32 masters, fp32 B=16 observe scan, losses and optimizer"),
        "config": {"workload": w["name"], "id": wid, "B_per_gpu": B, "T": T, "H": H, "global_batch": B * world,
                   "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": "256 MiB buffer written between timed steps (L2 flush)",
                   "cuda_graph": ("whole update = one graph replay" if use_graph else
                                  "five graph-replayed pieces cut at the NCCL collectives (dv3_dp_piece)" if world > 1 and not args.no_graph
                                  else False), "compute_mode": args.mode},
        "clocks": clk,
        "e2e": {"value": world * 1e3 / ms_e2e, "unit": "updates/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": int(launches),
        "scan_step_latency_us": scan_us,
        "step_tflops": gflop / ms,  # algorithmic GFLOP per update / ms = TFLOP/s
        "final_losses": final,
        "wall_s_timed_region": wall,
    }
    if replay_leg is not None:
        line["e2e_device_replay"] = replay_leg
    line["roofline"] = measure_roofline(e, cfg, pk, args)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(wid)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_roofline(e, cfg, pk, args):
    """Dominant kernel class of the update = the contractions over the (H+1)*N dream rows (heads forward / backward).
    Timed live with CUDA events on the launching stream, same kernel and shape as inside the step: the first-layer head
    GEMM [R, G+Z] x [G+Z, D]. In tensor-core mode that is the TMA-fed tcgen05 kernel on the bf16 twin of [h, z] (written by
    the producer kernels) and the bf16 weight shadow. The launches rotate over operand sets larger than L2 together, and
    are replayed from a CUDA graph so that the Python launch overhead is not in the per-launch time."""
    from dreamer_v3_b200 import ops
    R, K, Nn = (cfg.H + 1) * cfg.N, cfg.G + cfg.Z, cfg.D
    Bm = torch.randn(K, Nn, device="cuda")
    a_bytes = R * K * (2 if args.mode == 1 else 4)
    nset = max(2, int(np.ceil(160e6 / (a_bytes + 4.0 * R * Nn))))       # > 126 MB L2 across the rotation
    runs = []
    for _ in range(nset):
        A = torch.randn(R, K, device="cuda")
        C = torch.empty(R, Nn, device="cuda")
        runs.append(ops.gemm_bf16(ops.to_bf16(A), Bm, C_=C) if args.mode == 1 else (lambda A=A, C=C: ops.gemm(A, Bm, C_=C, mode=0)))
        del A
    reps = 5 * nset
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for r in runs:
            r()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(reps):
                runs[i % nset]()
        gr.replay()
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        gr.replay()
        t.record()
        torch.cuda.synchronize()
    ms = s.elapsed_time(t) / reps
    flops = 2.0 * R * K * Nn
    alg_bytes = float(a_bytes) + (2.0 if args.mode == 1 else 4.0) * K * Nn + 4.0 * R * Nn
    tf = flops / (ms * 1e-3) / 1e12
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    # the roofline that binds this shape: with bf16 operands the head contraction has ~180 flop/B, below the ridge
    # (measured 1634.5 TF / 6.49 TB/s = 252 flop/B), so it is HBM-bound; the fp32 SIMT variant is far below both
    hbm_bound = (alg_bytes / (pk["hbm"] * 1e9)) > (flops / (pk["tf"] * 1e12))
    out = {"kernel": f"gemm [{R}x{K}]x[{K}x{Nn}] ({'fp32 SIMT' if args.mode == 0 else 'bf16 tcgen05, TMA-fed (gemm_tma_kernel)'})",
           "traffic": TRAFFIC_BYTES.get((R, K, Nn, args.mode)), "algorithmic_bytes": alg_bytes, "algorithmic_flops": flops,
           "peak_source": pk["src"] + " (burst, kernel timed alone)", "ms_per_launch": ms,
           "l2": f"{nset} operand sets rotated ({nset * (a_bytes + 4 * R * Nn) / 1e6:.0f} MB > L2)", "launches_timed": reps,
           "tensor_tflops": tf, "tensor_frac": tf / pk["tf"], "hbm_gbs": gbs, "hbm_frac": gbs / pk["hbm"]}
    if hbm_bound:
        out.update({"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"]})
    else:
        out.update({"bound": "tensor", "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"]})
    return out


def cpu_baseline(wid):
    """Oracle restatement of the reference algorithm timed on this box's host cores (bounded sample)."""
    from dreamer_v3_b200.spec import init_params
    from oracle.dreamer_oracle import Oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = {"c1": 16, "c2": 16, "c3": 2, "c4": 1, "L": 1, "c5": 1}[wid]
    cfg = make_cfg(wid, B=rows)
    sample = synthetic_sample(cfg, 0)
    rng = np.random.default_rng(0)
    o = Oracle(cfg, init_params(cfg, seed=0), dtype=torch.float32)

    def step():
        res = o.train_world_model_one_step(sample, rng=rng)
        f = res["WORLD_MODEL_forward_train_outs"]
        o.train_actor_and_critic_one_step(f["h_states_BxT"], f["z_states_BxT"], sample["is_terminated"], H=cfg.H, rng=rng)

    step()
    n, t0 = 0, time.perf_counter()
    while n < 2 or (time.perf_counter() - t0 < 10.0 and n < 20):
        step(); n += 1
    dt = (time.perf_counter() - t0) / n * (16.0 / rows)
    return {"value": 1.0 / dt, "unit": "updates/s", "cores": cores, "kind": "port",
            "sample": f"{n} full updates at B={rows} of 16 rows (x{16 // rows} scaled), T=64, H=15, PyTorch-CPU fp32, {cores} threads"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--mode", type=int, default=1, help="1 = bf16 tcgen05 tensor cores for contractions with >= 128 rows (fp32 accumulate, fp32 scan / elementwise); 0 = fp32 everywhere")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-rows", type=int, default=None)
    ap.add_argument("--breakdown", action="store_true", help="print per-phase device times (ms) and exit")
    ap.add_argument("--quick", action="store_true", help="profiling aid: only the device-resident leg, no minimum warm-up")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
