"""2-rank data-parallel parity on GPUs (run under torchrun): sharding the replay batch over ranks + NCCL all-reduce /
all-gather must reproduce the single-process gradients on the full batch.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_parity.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dreamer_v3_b200.models.components import MLP, VectorDecoder  # noqa: E402
from dreamer_v3_b200.models.dreamer_model import DreamerModel  # noqa: E402
from dreamer_v3_b200.models.world_model import WorldModel  # noqa: E402
from dreamer_v3_b200.spec import Box, Discrete, param_spec  # noqa: E402
from dreamer_v3_b200.training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step  # noqa: E402
from oracle.dreamer_oracle import make_synthetic_sample  # noqa: E402  (data generator only)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, T, H = 4, 8, 5
Bl = B // world
ok = True
for discrete in (True, False):
    space = Discrete(3) if discrete else Box(-1, 1, (2,))

    def build(b, ws, rk):
        wm = WorldModel(model_dimension="XXS", action_space=space, batch_length_T=T, encoder=MLP(model_dimension="XXS"),
                        decoder=VectorDecoder(model_dimension="XXS", observation_space=Box(-1, 1, (5,))), seed=7)
        dm = DreamerModel(model_dimension="XXS", action_space=space, batch_size_B=b, batch_length_T=T, horizon_H=H,
                          world_model=wm, world_size=ws, rank=rk)
        g = torch.Generator().manual_seed(0)
        for n in ("rew.head.w", "critic.head.w"):
            dm.engine.tensor(n).copy_(0.02 * torch.randn(dm.engine.tensor(n).shape, generator=g))
        dm.critic.init_ema()
        return wm, dm

    wm, dm = build(Bl, world, rank)
    cfg_full = None
    rng = np.random.default_rng(3)
    Zc = dm.engine.cfg.Zc
    N = B * T
    u_post = rng.random((T, B, Zc), dtype=np.float32)
    u_dyn = rng.random((H, N, Zc), dtype=np.float32)
    act_noise = rng.random((H + 1, N), dtype=np.float32) if discrete else rng.standard_normal((H + 1, N, 2)).astype(np.float32)

    class C:  # full-batch config for the data generator
        pass
    from dreamer_v3_b200.spec import EngineConfig
    full_cfg = EngineConfig.make("XXS", action_space=space, obs_dim=5, B=B, T=T, H=H)
    sample = make_synthetic_sample(full_cfg, seed=5)
    rows = slice(rank * Bl, (rank + 1) * Bl)
    nrows = slice(rank * Bl * T, (rank + 1) * Bl * T)
    shard = {k: v[rows] for k, v in sample.items()}

    def step(wm_, dm_, smp, b, up, ud, an):
        wm_.set_noise(u_post=up, u_dyn=ud, act_noise=an)
        r = train_world_model_one_step(sample=smp, batch_size_B=b, batch_length_T=T, grad_clip=1000.0, world_model=wm_)
        a = train_actor_and_critic_one_step(
            world_model_forward_train_outs=r["WORLD_MODEL_forward_train_outs"],
            is_terminated=torch.as_tensor(smp["is_terminated"]).reshape(-1), horizon_H=H, gamma=0.997, lambda_=0.95,
            actor_grad_clip=100.0, critic_grad_clip=100.0, disagree_grad_clip=100.0, dreamer_model=dm_,
            entropy_scale=3e-4, return_normalization_decay=0.99)
        torch.cuda.synchronize()
        return r, a

    step(wm, dm, shard, Bl, u_post[:, rows], u_dyn[:, nrows], act_noise[:, nrows])
    e = dm.engine
    grads = {pi.name: e.tensor("grad." + pi.name).clone() for pi in param_spec(e.cfg) if pi.group != "critic_ema"}
    pct = e.tensor("ac.pct_ema").clone()
    if rank == 0:
        # single-process reference on the full batch (world_size 1 engine on the same GPU)
        dist_initialized = True
        wm1, dm1 = build(B, 1, 0)
        # bypass the DP branch: world_size == 1 on this engine
        step(wm1, dm1, sample, B, u_post, u_dyn, act_noise)
        e1 = dm1.engine
        worst = 0.0
        for n, g_dp in grads.items():
            g1 = e1.tensor("grad." + n)
            err = float((g_dp - g1).abs().max() / (g1.abs().max() + 1e-12))
            worst = max(worst, err)
            if err > 1e-4:
                ok = False
                print(f"  MISMATCH {n}: rel err {err:.3e}")
        perr = float((pct - e1.tensor("ac.pct_ema")).abs().max())
        print(f"{'discrete' if discrete else 'box'}: worst DP-vs-single gradient rel err {worst:.3e}, percentile EMA abs err {perr:.3e}")
        ok = ok and perr < 1e-5
        e1.close()
    dist.barrier()
    e.close()
if rank == 0:
    print("DP PARITY", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
