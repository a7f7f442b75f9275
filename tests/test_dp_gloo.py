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


# ---- the control flow of the package's own data-parallel update functions on gloo: which pieces run, which collectives on
# ---- which gradient spans, in which order - with the engine replaced by a recorder over CPU tensors (the GPU run of the same
# ---- code is tools/dp_parity.py / tests/test_gpu_dp.py)
class _StubEngine:
    """Quacks like dreamer_v3_b200.engine.Engine for training/train_one_step.py: three contiguous world-model gradient spans
    (encoder | recurrent | heads) in one flat buffer, actor + critic in another; dp_piece records and 'computes' gradients."""

    def __init__(self, rank, world):
        from dreamer_v3_b200.spec import EngineConfig
        self.cfg = EngineConfig.make("nano", B=2, T=4, H=2, A=2, discrete=True)
        self.world_size, self.rank = world, rank
        self.wm = torch.zeros(10 + 20 + 30)
        self.ac = torch.zeros(8 + 6)
        self.calls = []
        self.hp = type("HP", (), {})()
        self._t = {}

    def group_buffer(self, group, kind="grad"):
        assert kind == "grad"
        spans = {"world_model": self.wm, "world_model_encoder": self.wm[:10], "world_model_recurrent": self.wm[10:30],
                 "world_model_heads": self.wm[30:], "actor_critic": self.ac, "critic": self.ac[8:]}
        self.calls.append(("buffer", group))
        return spans[group]

    def dp_piece(self, piece, seed=0):
        self.calls.append(("piece", piece))
        r = float(self.rank + 1)
        if piece in (0, 5):
            self.wm.zero_()
        if piece in (0, 5):
            self.wm[30:] += r            # heads: final after the head backward / reverse-time scan
        if piece in (0, 6):
            self.wm[10:30] += 10 * r     # recurrent weight gradients
        if piece in (0, 7):
            self.wm[:10] += 100 * r      # encoder backward
        if piece == 3:
            self.ac[:] = r
        if piece == 1:
            self.at_adam = self.wm.clone()   # what the optimizer step sees

    def set_sample(self, sample):
        pass

    def write(self, name, value):
        pass

    def tensor(self, name):
        if name not in self._t:
            n = self.cfg.H * self.cfg.N
            self._t[name] = (torch.full((n,), float(self.rank)) if name == "ac.value_targets"
                             else torch.zeros(self.world_size * n) if name == "ac.value_targets_all" else torch.zeros(16))
        return self._t[name]


def _flow_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dreamer_v3_b200.training import train_one_step as tos
    opt = type("Opt", (), {"learning_rate": 1e-4, "epsilon": 1e-8})()
    for overlap in ("0", "1"):
        os.environ["DV3_DP_OVERLAP"] = overlap
        e = _StubEngine(rank, world)
        wm = type("WM", (), {})()
        wm.engine, wm.optimizer, wm._explicit_noise, wm.noise_seed = e, opt, True, 0
        wm._configure = lambda **kw: None
        wm._forward_outs = lambda: {"h_states_BxT": e.tensor("wm.hz")}
        wm.initial_h = None
        tos.train_world_model_one_step(sample={}, batch_size_B=2, batch_length_T=4, grad_clip=1000.0, world_model=wm)
        tot = sum(r + 1.0 for r in range(world))
        want = torch.cat([torch.full((10,), 100 * tot), torch.full((20,), 10 * tot), torch.full((30,), tot)])
        ok = bool(torch.equal(e.at_adam, want))      # every span reduced exactly once, and before the optimizer piece
        pieces = [c[1] for c in e.calls if c[0] == "piece"]
        bufs = [c[1] for c in e.calls if c[0] == "buffer"]
        # actor / critic half: value targets all-gathered between pieces 2 and 3, one collective for both gradient buffers
        dm = type("DM", (), {})()
        dm.engine, dm.world_model = e, wm
        dm.actor = type("A", (), {"optimizer": opt})()
        dm.critic = type("C", (), {"optimizer": opt, "ema_decay": 0.98})()
        dm._dream_outs = lambda: {}
        e.calls.clear()
        hz = e.tensor("wm.hz")
        real_is_cuda = torch.Tensor.is_cuda
        try:   # the function takes its graph-piece route when it is handed the engine's own (CUDA) wm.hz view
            torch.Tensor.is_cuda = property(lambda self: True)
            tos.train_actor_and_critic_one_step(
                world_model_forward_train_outs={"h_states_BxT": hz, "z_states_BxT": hz}, is_terminated=torch.zeros(8), horizon_H=2,
                gamma=0.997, lambda_=0.95, actor_grad_clip=100.0, critic_grad_clip=100.0, disagree_grad_clip=100.0, dreamer_model=dm,
                entropy_scale=3e-4, return_normalization_decay=0.99)
        finally:
            torch.Tensor.is_cuda = real_is_cuda
        ac_pieces = [c[1] for c in e.calls if c[0] == "piece"]
        gathered = e.tensor("ac.value_targets_all")
        n = e.cfg.H * e.cfg.N
        ok_ac = bool(torch.equal(e.ac, torch.full((14,), tot))) and all(
            bool((gathered[r * n:(r + 1) * n] == float(r)).all()) for r in range(world))
        if rank == 0:
            out[overlap] = dict(ok=ok, pieces=pieces, bufs=bufs, ok_ac=ok_ac, ac_pieces=ac_pieces)
    dist.barrier()
    dist.destroy_process_group()


def test_update_functions_issue_the_right_collectives_on_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_flow_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    plain, over = out["0"], out["1"]
    assert plain["ok"] and plain["pieces"] == [0, 1] and plain["bufs"] == ["world_model"]
    # overlapped: heads bucket after piece 5, recurrent bucket after piece 6, encoder bucket after piece 7, then the optimizer
    assert over["ok"] and over["pieces"] == [5, 6, 7, 1]
    assert over["bufs"] == ["world_model_heads", "world_model_recurrent", "world_model_encoder"]
    for r in (plain, over):
        assert r["ok_ac"] and r["ac_pieces"] == [2, 3, 4]
