"""Data-parallel contract on CPU (gloo, world_size 2): sharding the replay batch over ranks and all-reducing
(sum) gradients that each rank normalised by the GLOBAL batch reproduces the single-process gradients, and
all-gathering the value targets reproduces the global percentiles. The oracle stands
in for the engine on CPU; the collective calls are the ones train_one_step.py issues on NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dreamer_v3_b200.spec import EngineConfig, init_params
from oracle import numerics as nx
from oracle.dreamer_oracle import Oracle, make_synthetic_sample


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    full = EngineConfig.make("nano", B=4, T=5, H=3, A=2, discrete=True)
    P = init_params(full, seed=0, mode="random")
    sample = make_synthetic_sample(full, seed=1)
    rng = np.random.default_rng(2)
    u_post = rng.random((full.T, full.B, full.Zc), dtype=np.float32)
    rows = slice(rank * full.B // world, (rank + 1) * full.B // world)
    local = EngineConfig.make("nano", B=full.B // world, T=5, H=3, A=2, discrete=True)
    o = Oracle(local, P)
    shard = {k: v[rows] for k, v in sample.items()}
    res = o.train_world_model_one_step(shard, u_post=u_post[:, rows], apply=False)
    names = o.group_names("world_model")
    flat = torch.cat([res["grads"][n].reshape(-1) for n in names]) / world   # local mean -> global mean share
    dist.all_reduce(flat)                                                   # what train_one_step.py does on NCCL
    # value-target all-gather for the percentile
    vt = torch.arange(6, dtype=torch.float32).reshape(3, 2) + 10 * rank
    gathered = torch.empty(world * vt.numel())
    dist.all_gather_into_tensor(gathered, vt.reshape(-1))
    if rank == 0:
        o_full = Oracle(full, P)
        ref = o_full.train_world_model_one_step(sample, u_post=u_post, apply=False)
        ref_flat = torch.cat([ref["grads"][n].reshape(-1) for n in names])
        out["grad_err"] = float((flat - ref_flat).abs().max() / ref_flat.abs().max())
        allv = np.concatenate([np.arange(6, dtype=np.float32) + 10 * r for r in range(world)])
        out["pct_ok"] = bool(nx.percentile_nearest(gathered.numpy(), 95) == nx.percentile_nearest(allv, 95))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_match_full_batch():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["grad_err"] < 1e-5, dict(out)
    assert out["pct_ok"]
