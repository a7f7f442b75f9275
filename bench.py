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
def algorithmic_gflop(cfg):
    """Algorithmic GFLOP of one train update derived from the layer shapes of `cfg` (SURVEY.md 8d rules: FLOP = 2 MAC;
    backward = 2x forward where a weight gradient is needed, 1x where only the input gradient is needed; the Discrete actor
    has no BPTT through the dream, the Box actor has). Reproduces the SURVEY table (150 / 235 / 827 / 2319 / 3506 / 8756)."""
    G, D, L, Z, A, E, m = cfg.G, cfg.D, cfg.L, cfg.Z, cfg.A, cfg.E, cfg.cnn_mult
    N, H = cfg.N, cfg.H
    R = (H + 1) * N
    mlp = lambda k: k * D + (L - 1) * D * D          # noqa: E731  MACs per row of an L-layer MLP on k inputs
    if cfg.image_obs:
        enc, cin, side = 0, cfg.img_c, cfg.img_hw
        for l in range(4):
            cout, side = m << l, side // 2
            enc += side * side * 16 * cin * cout
            cin = cout
        dec, cin, side = (G + Z) * E, 8 * m, 4
        for l in range(3):
            dec += side * side * 16 * cin * (cin // 2)
            cin, side = cin // 2, side * 2
        dec += side * side * 16 * cin * cfg.img_c
    else:
        enc = mlp(cfg.obs_dim)
        dec = mlp(G + Z) + D * cfg.obs_dim
    step = (Z + A) * D + D * 3 * G + G * 3 * G + (E + G) * D + 2 * D * Z + G * D   # one observe-scan step, per row
    rew, cont = mlp(G + Z) + D * 255, mlp(G + Z) + D
    wm = 3 * N * (enc + step + dec + rew + cont)
    seq_dyn = (Z + A) * D + (D + G) * 3 * G + G * D + D * Z
    actor = (mlp(G + Z) + D * A) * (1 if cfg.discrete else 2)
    critic = mlp(G + Z) + D * 255
    ac = H * N * seq_dyn + R * (actor + rew + cont + 2 * critic) + 2 * H * N * (critic + actor)
    if not cfg.discrete:
        ac += H * N * seq_dyn + R * (rew + critic)
    return {"world_model": 2e-9 * wm, "actor_critic": 2e-9 * ac, "total": 2e-9 * (wm + ac), "observe_scan_fwd": 2e-9 * N * step}


def scan_step_weight_bytes(cfg, bytes_per_weight=4):
    """Weights one observe-scan step has to read that are on the recurrence (the step's weights minus the
    hoisted encoder / action parts): W_pre[:Z], GRU kernel + recurrent kernel, W_post[E:], posterior representation layer.
    (The prior net is hoisted out of the loop entirely.)"""
    G, D, Z = cfg.G, cfg.D, cfg.Z
    return bytes_per_weight * (Z * D + D * 3 * G + G * 3 * G + G * D + D * Z)


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the head-contraction kernel from the committed ncu --set full
# capture (profiles/r01_gemm_tma_ncu.txt; the fp32 C tile stays in L2 past kernel end), keyed by (M, K, N, mode)
TRAFFIC_BYTES = {(16384, 1280, 256, 1): 42616320 + 142080}


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


def workload_config(wid, world):
    """The `config` block of the JSON line: the workload both arms are measured on (the reference arm runs "on your arm's
    config", so the two lines carry the same dict; what describes an implementation - graph replay, compute mode - is in the
    top-level `engine` block of the GPU arm)."""
    return {"workload": WORKLOADS[wid]["name"], "id": wid, "B_per_gpu": 16, "T": 64, "H": 15, "global_batch": 16 * world,
            "parallelism": f"dp{world}" if world > 1 else "single",
            "l2": "GPU arm: 256 MiB buffer written between timed steps (L2 flush); not applicable to the CPU reference arm"}


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
        "config": workload_config(wid, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": "port", "sample": samp},
        "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm, PyTorch-CPU restatement (TensorFlow unavailable in this image)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- our arm
class Ctx:
    """Process-wide state of the benchmarked arm: ranks, timing helpers."""

    def __init__(self, args):
        import torch.distributed as dist
        self.args, self.dist = args, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, steps):
        evs = []
        for _ in range(steps):
            self.flush.zero_()  # L2 flush (256 MiB > 126 MB L2) between timed iterations, outside the timed events
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); t.record()
            evs.append((s, t))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / steps

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


class Job:
    """One workload built through the reference-facing API (WorldModel / DreamerModel mirrors) on this rank."""

    def __init__(self, ctx, wid, B_local, mode):
        from dreamer_v3_b200.models.components import CNNAtari, ConvTransposeAtari, MLP, VectorDecoder
        from dreamer_v3_b200.models.dreamer_model import DreamerModel
        from dreamer_v3_b200.models.world_model import WorldModel
        from dreamer_v3_b200.spec import Box, Discrete
        self.ctx, self.wid, self.w = ctx, wid, WORKLOADS[wid]
        w = self.w
        self.cfg = cfg = make_cfg(wid, B=B_local)
        aspace = Discrete(w["A"]) if w["discrete"] else Box(-1.0, 1.0, (w["A"],))
        size = w["size"]
        if cfg.image_obs:
            enc, dec = CNNAtari(model_dimension=size), ConvTransposeAtari(model_dimension=size, gray_scaled=False)
        else:
            enc, dec = MLP(model_dimension=size), VectorDecoder(model_dimension=size, observation_space=Box(-1, 1, (cfg.obs_dim,)))
        self.wm = WorldModel(model_dimension=size, action_space=aspace, batch_length_T=cfg.T, encoder=enc, decoder=dec,
                             symlog_obs=not cfg.image_obs, compute_mode=mode)
        self.dm = DreamerModel(model_dimension=size, action_space=aspace, batch_size_B=cfg.B, batch_length_T=cfg.T,
                               horizon_H=cfg.H, world_model=self.wm, world_size=ctx.world, rank=ctx.rank)
        self.e = self.dm.engine
        self.dm.critic.init_ema()
        sample_np = synthetic_sample(cfg, seed=ctx.rank)
        self.pinned = {k: torch.from_numpy(v).pin_memory() for k, v in sample_np.items()}
        self.h2d_bytes = sum(v.numel() * 4 for v in self.pinned.values())
        self.host_scalars = torch.empty(24, dtype=torch.float32).pin_memory()
        self.d2h_bytes = self.host_scalars.numel() * 4
        self.hp = dict(horizon_H=cfg.H, gamma=0.997, lambda_=0.95, actor_grad_clip=100.0, critic_grad_clip=100.0,
                       disagree_grad_clip=100.0, entropy_scale=3e-4, return_normalization_decay=0.99)
        self.use_graph = ctx.world == 1 and not ctx.args.no_graph
        self.e.set_sample(self.pinned)

    def close(self):
        self.e.close()
        self.wm = self.dm = self.e = None
        torch.cuda.empty_cache()

    def api_step(self, host_inputs):
        from dreamer_v3_b200.training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step
        e, cfg = self.e, self.cfg
        sample = self.pinned if host_inputs else {}   # {}: inputs already resident in the engine's in.* buffers
        r = train_world_model_one_step(sample=sample, batch_size_B=cfg.B, batch_length_T=cfg.T, grad_clip=1000.0, world_model=self.wm)
        a = train_actor_and_critic_one_step(world_model_forward_train_outs=r["WORLD_MODEL_forward_train_outs"],
                                            is_terminated=e.tensor("in.is_terminated").reshape(-1), dreamer_model=self.dm, **self.hp)
        if host_inputs:
            self.host_scalars[:8].copy_(e.tensor("wm.scalars"), non_blocking=True)
            self.host_scalars[8:24].copy_(e.tensor("ac.scalars"), non_blocking=True)
        return r, a

    def device_step(self):
        if self.use_graph:
            self.e.train_update(noise_seed=1234, use_graph=True)
        else:
            self.api_step(False)

    def e2e_step(self):
        self.api_step(True)
        torch.cuda.current_stream().synchronize()  # the caller reads the losses

    def measure(self, steps, warmup, clocks=False, e2e=True, phases=True):
        """device-resident leg (+ clocks), end-to-end leg, scan latency and the per-phase device times of one update."""
        ctx, e, cfg = self.ctx, self.e, self.cfg
        out = {}
        for _ in range(max(warmup, 3)):
            self.device_step()
        ctx.barrier()
        sampler = ClockSampler(ctx.local) if clocks and ctx.rank == 0 else None
        if sampler:
            sampler.start()
        l0 = e.kernel_launch_count()
        wall0 = time.perf_counter()
        ms = ctx.timed(self.device_step, steps)
        ctx.barrier()
        out["wall_s_timed_region"] = time.perf_counter() - wall0
        out["gpu_launches"] = int(e.kernel_launch_count() - l0)
        if sampler:
            out["clocks"] = sampler.stop()
        out["ms_per_step"] = ms = ctx.max_over_ranks(ms)
        if e2e:
            for _ in range(3):
                self.api_step(True)
            ctx.barrier()
            ms_e2e = ctx.max_over_ranks(ctx.timed(self.e2e_step, steps))
            ctx.barrier()
            out["e2e_ms_per_step"] = ms_e2e
            hs = self.host_scalars
            out["final_losses"] = {"wm_total": float(hs[6]), "critic_total": float(hs[8]), "actor_total": float(hs[11])}
        if phases:
            # per-phase device times of the non-graph path (CUDA events at the labelled phase boundaries, dv3_phase_report);
            # "obs.scan" = the observe-scan kernel(s) alone, "wmb.scan_bwd" = the reverse-time BPTT scan alone
            e.fill_noise(1, 0, 3)
            reps = 3
            for i in range(reps + 1):
                if i == 1:
                    e.phase_report(enable=1)      # discard the warm-up pass
                e.observe_forward(); e.wm_loss_backward()
                e.dream_forward(start_from_observe=True); e.ac_loss_backward(3)
            ph = e.phase_report(enable=0)
            out["phases_ms"] = {k: round(v[0] / reps, 4) for k, v in ph.items()}
            T = cfg.T
            out["scan_fwd_us_per_step"] = ctx.max_over_ranks(ph.get("obs.scan", (0.0, 1))[0] / reps) * 1e3 / T
            out["scan_bwd_us_per_step"] = ctx.max_over_ranks(ph.get("wmb.scan_bwd", (0.0, 1))[0] / reps) * 1e3 / T
        return out


# dram__bytes_read.sum + dram__bytes_write.sum per timestep of the scan kernels from the committed ncu --set full captures
# (profiles/r02_scan_tc_xl_ncu.txt: XL forward 8798.9 + 197.7 MB, backward 8977.4 + 173.1 MB over T = 64 steps;
#  profiles/r02_c3_kernels_ncu_table.txt: S backward 46.8 + 5.6 MB - the bf16 weights stay in the L2)
SCAN_TRAFFIC_PER_STEP = {("XL", "fwd"): (8798.88e6 + 197.66e6) / 64, ("XL", "bwd"): (8977.40e6 + 173.10e6) / 64,
                         ("S", "bwd"): (46.80e6 + 5.60e6) / 64}


def scan_roofline(cfg, us_fwd, us_bwd, pk, mode):
    """Roofline entry of the time-dominant latency / bandwidth kernel of an update: the observe scan. Algorithmic bytes per
    step = the step's weights on the recurrence, read once: bf16 in the tensor-core mode at S ... XL (the
    weight-streaming tcgen05 kernels of dv3_scan_tc.cu; the backward kernel also streams W_pre[:Z]), fp32 otherwise. At XS the
    weights are resident in the cluster's shared memory for all T steps, so the kernel has no HBM roofline: a latency chain."""
    G, D, Z = cfg.G, cfg.D, cfg.Z
    tc = mode == 1 and G >= 512 and D % 128 == 0 and G % 128 == 0 and cfg.B <= 16
    bpw = 2 if tc else 4
    fwd_bytes = bpw * (D * 3 * G + G * 3 * G + G * D + D * Z) + (0 if tc else bpw * Z * D)
    bwd_bytes = bpw * (D * 3 * G + G * 3 * G + G * D + D * Z + Z * D)
    size = cfg.model_dimension
    out = {"kernel": ("scan_tc_kernel<FwdPolicy> / <BwdPolicy> (bf16 weight stream, swap-AB tcgen05)" if tc else
                      "observe scan forward / backward"),
           "weight_bytes_per_step": fwd_bytes, "bwd_weight_bytes_per_step": bwd_bytes, "weight_dtype": "bf16" if tc else "f32",
           "us_per_step": us_fwd, "bwd_us_per_step": us_bwd, "peak": pk["hbm"], "unit": "GB/s", "peak_source": pk["src"],
           "traffic": SCAN_TRAFFIC_PER_STEP.get((size, "fwd")), "bwd_traffic": SCAN_TRAFFIC_PER_STEP.get((size, "bwd"))}
    if cfg.G <= 256:
        if mode == 1:
            out["kernel"] = "observe_scan_cluster_kernel<true> / observe_scan_bwd_cluster_kernel (bf16 weights resident, mma.sync)"
            out["weight_dtype"] = "bf16 (resident in shared memory)"
            note = ("weights resident in the 16-CTA cluster's shared memory in bf16 for all T steps (no weight bytes from HBM / L2 per "
                    "step; the one-hot row gather of W_pre included): per step 4 dependent 16-row mma.sync stages, each followed by a "
                    "bulk-copy exchange through distributed shared memory signalled on the receiver's mbarrier (no cluster barrier)")
        else:
            note = ("weights resident in the 16-CTA cluster's shared memory for all T steps (no weight bytes from HBM / L2 per step): "
                    "4 dependent 16-row contraction stages + exchanges per step (backward: two-stage grid-wide kernel)")
        out.update({"bound": "latency", "note": note, "achieved": None, "frac": None})
    else:
        gbs = fwd_bytes / (us_fwd * 1e-6) / 1e9 if us_fwd > 0 else None
        gbs_b = bwd_bytes / (us_bwd * 1e-6) / 1e9 if us_bwd > 0 else None
        out.update({"bound": "hbm", "achieved": gbs, "frac": gbs / pk["hbm"] if gbs else None,
                    "bwd_achieved": gbs_b, "bwd_frac": gbs_b / pk["hbm"] if gbs_b else None,
                    "note": ("the step's weights are re-read every step: from HBM at XL (136 MB bf16 > what the 126 MB L2 keeps), from the "
                             "L2 at S ... L, where the step is a chain of 4 - 5 grid-wide stages (barrier + staging latency), not a "
                             "bandwidth limit")})
    return out


def allreduce_ms(ctx, job):
    """Device time of the gradient all-reduces of one update in isolation (NCCL over NVLink), max over ranks."""
    e, dist = job.e, ctx.dist
    bufs = [e.group_buffer(g, "grad") for g in ("world_model", "actor_critic")]   # the two collectives an update issues
    for _ in range(2):
        for b in bufs:
            dist.all_reduce(b)
    ctx.barrier()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    s.record()
    for _ in range(reps):
        for b in bufs:
            dist.all_reduce(b)
    t.record()
    torch.cuda.synchronize()
    return ctx.max_over_ranks(s.elapsed_time(t) / reps), sum(b.numel() * 4 for b in bufs)


def config_entry(ctx, job, m, pk, mode):
    cfg = job.cfg
    gf = algorithmic_gflop(cfg)
    ent = {"workload": job.w["name"], "B_per_gpu": cfg.B, "ms_per_step": m["ms_per_step"], "gpu_launches": m["gpu_launches"],
           "algorithmic_gflop": round(gf["total"], 1), "step_tflops": gf["total"] / m["ms_per_step"],
           "tensor_frac_of_sustained_peak": gf["total"] / m["ms_per_step"] / pk["tf_sustained"]}
    if "e2e_ms_per_step" in m:
        ent["e2e_ms_per_step"] = m["e2e_ms_per_step"]
        ent["h2d_bytes_per_step"], ent["d2h_bytes_per_step"] = job.h2d_bytes, job.d2h_bytes
    if "phases_ms" in m:
        ent["scan_step_latency_us"] = m["scan_fwd_us_per_step"]
        ent["scan_bwd_step_latency_us"] = m["scan_bwd_us_per_step"]
        ent["phases_ms"] = m["phases_ms"]
        ent["roofline"] = scan_roofline(cfg, m["scan_fwd_us_per_step"], m["scan_bwd_us_per_step"], pk, mode)
    return ent


def run_ours(args):
    ctx = Ctx(args)
    world, rank = ctx.world, ctx.rank
    dist = ctx.dist
    pk = peaks()
    wid = args.workload
    job = Job(ctx, wid, 16, args.mode)
    e, cfg = job.e, job.cfg
    B, T, H = cfg.B, cfg.T, cfg.H

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
    if args.quick:
        m = job.measure(args.steps, args.warmup, e2e=False, phases=False)
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": m["ms_per_step"], "value": world * 1e3 / m["ms_per_step"],
                              "gpu_launches": m["gpu_launches"]}), flush=True)
        return

    m = job.measure(args.steps, args.warmup, clocks=True)
    ms, ms_e2e = m["ms_per_step"], m["e2e_ms_per_step"]
    value = world * 1e3 / ms

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
            job.api_step(False)
            job.host_scalars[:8].copy_(e.tensor("wm.scalars"), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3):
            replay_step()
        ms_rp = ctx.timed(replay_step, args.steps)
        replay_leg = {"value": 1e3 / ms_rp, "unit": "updates/s", "ms_per_step": ms_rp,
                      "what": "EpisodeReplayBuffer.sample (device store, dv3_replay_gather) + both update steps + D2H of the losses"}
        e.set_sample(job.pinned)

    # ---- whole observe_forward / T (encoder and heads included), kept for continuity with round 1 ----
    for _ in range(3):
        e.observe_forward()
    obs_us = ctx.max_over_ranks(ctx.timed(e.observe_forward, max(5, args.steps // 2))) * 1e3 / T

    ar = allreduce_ms(ctx, job) if world > 1 else None
    gf = algorithmic_gflop(cfg)
    line = {
        "metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == 0 else "bf16", "data": "synthetic",
        "precision_note": ("fp32 everywhere" if args.mode == 0 else "bf16 operands / fp32 accumulate on tcgen05 for contractions with >= 128 rows; fp32 masters, losses and optimizer"),
        "config": workload_config(wid, world),
        "engine": {"cuda_graph": ("whole update = one graph replay" if job.use_graph else
                                  "graph-replayed pieces cut at the NCCL collectives (dv3_dp_piece)" if world > 1 and not args.no_graph
                                  else False), "compute_mode": args.mode},
        "value_note": ("weak scaling: B=16 replay rows per GPU; value = B16xT64 batches consumed per second over all ranks "
                       "(= optimizer updates/s x n_gpus); literal updates/s of the sharded global batch are in strong_xl"),
        "clocks": m.get("clocks"),
        "e2e": {"value": world * 1e3 / ms_e2e, "unit": "updates/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": job.h2d_bytes, "d2h_bytes_per_step": job.d2h_bytes},
        "gpu_launches": m["gpu_launches"],
        "scan_step_latency_us": m["scan_fwd_us_per_step"],
        "scan_bwd_step_latency_us": m["scan_bwd_us_per_step"],
        "observe_forward_us_per_step": obs_us,
        "phases_ms": m["phases_ms"],
        "algorithmic_gflop": round(gf["total"], 1),
        "step_tflops": gf["total"] / ms,  # algorithmic GFLOP per update / ms = TFLOP/s
        "final_losses": m["final_losses"],
        "wall_s_timed_region": m["wall_s_timed_region"],
    }
    if ar is not None:
        line["allreduce"] = {"ms_per_update_isolated": ar[0], "bytes": ar[1], "busbw_gbs": 2 * (world - 1) / world * ar[1] / (ar[0] * 1e-3) / 1e9}
    if replay_leg is not None:
        line["e2e_device_replay"] = replay_leg
    line["roofline"] = measure_roofline(e, cfg, pk, args) if rank == 0 else None
    line["roofline_scan"] = scan_roofline(cfg, m["scan_fwd_us_per_step"], m["scan_bwd_us_per_step"], pk, args.mode)
    if args.mode == 1 and not args.no_fp32_line and world == 1:
        # the same update in the fp32 mode (compute_mode 0: the 1e-4 parity mode), device-resident leg only
        job.close()
        j0 = Job(ctx, wid, 16, 0)
        m0 = j0.measure(max(3, args.steps // 2), 3, e2e=False, phases=False)
        line["fp32_mode"] = {"value": 1e3 / m0["ms_per_step"], "ms_per_step": m0["ms_per_step"], "gpu_launches": m0["gpu_launches"]}
        j0.close()
    else:
        job.close()

    # ---- every other BASELINE configuration (the metric is "per model size") ----
    if not args.no_per_config:
        per = {}
        if world == 1:
            per[wid] = {"workload": job.w["name"], "value": value, "ms_per_step": ms, "e2e": line["e2e"]["value"],
                        "scan_step_latency_us": line["scan_step_latency_us"], "see": "top-level keys of this line"}
            for w2 in [k for k in WORKLOADS if k != wid]:
                j = Job(ctx, w2, 16, args.mode)
                k_steps = max(3, min(args.steps, 10 if w2 in ("c1", "c3") else 5))
                mm = j.measure(k_steps, 3)
                ent = config_entry(ctx, j, mm, pk, args.mode)
                ent["value"] = 1e3 / mm["ms_per_step"]
                ent["e2e"] = 1e3 / mm["e2e_ms_per_step"]
                ent["steps"] = k_steps
                per[w2] = ent
                j.close()
            line["per_config"] = per
        # BASELINE configs[4]: XL with the B=16 replay batch SHARDED over the ranks (strong scaling, literal updates/s)
        if world == 1 and "c5" in per and wid != "c5":
            x = per["c5"]
            line["strong_xl"] = {"n_gpus": 1, "B_global": 16, "B_per_gpu": 16, "value": x["value"], "unit": "updates/s",
                                 "ms_per_step": x["ms_per_step"], "scaling": "strong", "allreduce": None}
        elif world > 1 and 16 % world == 0:
            j = Job(ctx, "c5", 16 // world, args.mode)
            mm = j.measure(max(3, min(args.steps, 5)), 3, e2e=False)
            arx = allreduce_ms(ctx, j)
            line["strong_xl"] = {"n_gpus": world, "B_global": 16, "B_per_gpu": 16 // world, "value": 1e3 / mm["ms_per_step"],
                                 "unit": "updates/s", "ms_per_step": mm["ms_per_step"], "scaling": "strong",
                                 "phases_ms": mm.get("phases_ms"), "scan_step_latency_us": mm.get("scan_fwd_us_per_step"),
                                 "allreduce": {"ms_per_update_isolated": arx[0], "bytes": arx[1],
                                               "busbw_gbs": 2 * (world - 1) / world * arx[1] / (arx[0] * 1e-3) / 1e9}}
            j.close()
    if rank == 0:
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
    ap.add_argument("--no-per-config", action="store_true", help="only the headline workload (skip per_config / strong_xl)")
    ap.add_argument("--no-fp32-line", action="store_true", help="skip the fp32-mode value of the headline workload")
    ap.add_argument("--ref-rows", type=int, default=None)
    ap.add_argument("--breakdown", action="store_true", help="print per-phase device times (ms) and exit")
    ap.add_argument("--quick", action="store_true", help="profiling aid: only the device-resident leg, no minimum warm-up")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
