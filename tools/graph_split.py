"""Graph-mode device time of the two halves of an update (world-model step, actor/critic step), each replayed as its own CUDA
graph, for a bench workload and horizon: python tools/graph_split.py c2 [H ...]. The slope over H is the in-graph cost of one
dreamed step plus one BPTT step - the per-phase numbers of dv3_phase_report come from the non-graph path and are launch-bound."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def run(wid, H, steps=30):
    class A: no_graph = False
    ctx = bench.Ctx.__new__(bench.Ctx)
    ctx.world, ctx.rank, ctx.local, ctx.args, ctx.dist = 1, 0, 0, A, None
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    orig = bench.make_cfg
    bench.make_cfg = lambda w, B=16, T=64, H_=H: orig(w, B=B, T=T, H=H_)
    try:
        job = bench.Job(ctx, wid, 16, 1)
    finally:
        bench.make_cfg = orig
    e = job.e
    job.api_step(False)   # sets the hyper-parameter block the way the public API does
    for _ in range(5):
        e.train_world_model_step(noise_seed=7); e.train_actor_critic_step(noise_seed=7)
    torch.cuda.synchronize()
    tw = ta = 0.0
    for _ in range(steps):
        ctx.flush.zero_()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); e.train_world_model_step(noise_seed=7); b.record(); e.train_actor_critic_step(noise_seed=7); c.record()
        torch.cuda.synchronize()
        tw += a.elapsed_time(b); ta += b.elapsed_time(c)
    l0 = e.kernel_launch_count(); e.train_actor_critic_step(noise_seed=7); l1 = e.kernel_launch_count()
    out = {"workload": wid, "H": H, "wm_ms": round(tw / steps, 4), "ac_ms": round(ta / steps, 4), "ac_launches": l1 - l0,
           "env": {k: v for k, v in os.environ.items() if k.startswith("DV3_")}}
    job.close()
    return out


if __name__ == "__main__":
    wid = sys.argv[1] if len(sys.argv) > 1 else "c2"
    Hs = [int(x) for x in sys.argv[2:]] or [15]
    for H in Hs:
        print(json.dumps(run(wid, H)), flush=True)
