"""Error behaviour of the C ABI (include/dv3.h, SURVEY.md 8b "Errors"): every call returns 0 or a negative code with a
message in dv3_last_error, nothing throws across the ABI, arguments are validated before any launch, and a rejected call
leaves the engine usable. The reference raises Python exceptions / in-graph asserts at the same places (ranks and shapes
of the sample, `timesteps_H`, unknown model sizes); the Python mirror turns the codes into `Dv3Error`."""
import ctypes as C

import numpy as np
import pytest
import torch

from dreamer_v3_b200 import _lib
from dreamer_v3_b200.spec import EngineConfig

pytestmark = pytest.mark.gpu


def _config(**over):
    cfg = EngineConfig.make("nano", A=2, discrete=True, obs_dim=4, B=2, T=4, H=2)
    c = _lib.Dv3Config()
    c.abi_version = _lib.ABI_VERSION
    c.G, c.D, c.L, c.cnn_mult, c.Zc, c.Zk = cfg.G, cfg.D, cfg.L, cfg.cnn_mult, cfg.Zc, cfg.Zk
    c.A, c.discrete, c.image_obs, c.obs_dim = cfg.A, 1, 0, cfg.obs_dim
    c.img_hw, c.img_c, c.B, c.T, c.H = 64, 3, cfg.B, cfg.T, cfg.H
    c.symlog_obs, c.compute_mode, c.world_size, c.rank, c.device = 1, 0, 1, 0, torch.cuda.current_device()
    for k, v in over.items():
        setattr(c, k, v)
    return c


@pytest.mark.parametrize("over,needle", [
    (dict(abi_version=99), b"abi_version"),
    (dict(Zk=64), b"Zk"),
    (dict(D=4096), b"D must be"),
    (dict(G=0), b"non-positive"),
    (dict(T=-1), b"non-positive"),
    (dict(A=64), b"32 actions"),
    (dict(compute_mode=7), b"compute_mode"),
    (dict(image_obs=1, img_hw=32), b"64x64"),
    (dict(num_curiosity_nets=-1), b"num_curiosity_nets"),
    (dict(num_curiosity_nets=2, world_size=2), b"single-process"),
])
def test_create_rejects_bad_configs(over, needle):
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.dv3_create(C.byref(_config(**over)), C.byref(h))
    assert rc < 0 and not h.value
    assert needle in lib.dv3_last_error(None), lib.dv3_last_error(None)
    assert lib.dv3_create(None, C.byref(h)) < 0


def test_calls_validate_and_the_engine_survives():
    from dreamer_v3_b200.engine import Dv3Error, Engine
    from dreamer_v3_b200.spec import init_params
    from oracle.dreamer_oracle import make_synthetic_sample
    cfg = EngineConfig.make("nano", A=2, discrete=True, obs_dim=4, B=2, T=4, H=2)
    e = Engine(cfg)
    lib, h = e.lib, e.h
    try:
        e.load_params(init_params(cfg, seed=0))
        d = _lib.Dv3TensorDesc()
        assert lib.dv3_find_tensor(h, b"no.such.tensor", C.byref(d)) < 0
        assert lib.dv3_tensor_info(h, -1, C.byref(d)) < 0 and lib.dv3_tensor_info(h, 10 ** 6, C.byref(d)) < 0
        buf = np.zeros(3, np.float32)
        p = buf.ctypes.data_as(C.c_void_p)
        assert lib.dv3_write_tensor(h, b"no.such.tensor", p, 12, 0, None, 1) < 0 and b"unknown tensor" in lib.dv3_last_error(h)
        assert lib.dv3_write_tensor(h, b"in.rewards", p, 12, 0, None, 1) < 0 and b"size mismatch" in lib.dv3_last_error(h)
        assert lib.dv3_read_tensor(h, b"in.rewards", p, 12, 0, None, 1) < 0 and b"size mismatch" in lib.dv3_last_error(h)
        ptr, n = C.c_void_p(), C.c_int64()
        for group, kind in ((-1, 1), (9, 1), (3, 1), (0, 7), (5, 0), (5, 1)):   # 3 = EMA critic: parameters only; 5: curiosity is off
            assert lib.dv3_group_buffer(h, group, kind, C.byref(ptr), C.byref(n)) < 0, (group, kind)
        for group in (6, 7, 8):   # the three world-model gradient buckets tile the whole buffer
            assert lib.dv3_group_buffer(h, group, 1, C.byref(ptr), C.byref(n)) == 0 and n.value > 0
        total = 0
        for group in (6, 7, 8):
            lib.dv3_group_buffer(h, group, 1, C.byref(ptr), C.byref(n))
            total += n.value
        lib.dv3_group_buffer(h, 0, 1, C.byref(ptr), C.byref(n))
        assert total == n.value
        for group in (-1, 3, 4, 5, 6):
            assert lib.dv3_optimizer_step(h, group, 100.0, 1e-4, 1e-8, -1.0, None) < 0, group
        assert b"group" in lib.dv3_last_error(h)
        assert lib.dv3_forward_inference(h, 0, 0, None) < 0 and lib.dv3_forward_inference(h, cfg.B + 1, 0, None) < 0
        assert b"rows" in lib.dv3_last_error(h)
        assert lib.dv3_dp_piece(h, C.byref(e.hp), 8, 0, 0, None) < 0 and lib.dv3_dp_piece(h, C.byref(e.hp), -1, 0, 0, None) < 0
        assert lib.dv3_disagree_loss_backward(h, None) < 0 and b"curiosity" in lib.dv3_last_error(h)
        assert lib.dv3_set_inference_calls(h, -5) < 0
        with pytest.raises(Dv3Error):
            e.optimizer_step("critic_ema", 1.0, 1e-4, 1e-8)
        with pytest.raises(KeyError):
            e.load_params({"post.mlp0.w": np.zeros((1, 1), np.float32)})
        with pytest.raises(ValueError):
            bad = init_params(cfg, seed=0)
            bad["post.mlp0.w"] = bad["post.mlp0.w"][:-1]
            e.load_params(bad)
        # none of the rejected calls launched anything or broke the handle: a full update still runs and is finite
        e.set_sample(make_synthetic_sample(cfg, seed=1))
        e.train_update(noise_seed=3, use_graph=False)
        e.train_update(noise_seed=3, use_graph=True)
        torch.cuda.synchronize()
        assert torch.isfinite(e.tensor("wm.scalars")).all() and torch.isfinite(e.tensor("ac.scalars")[:7]).all()
        assert int(e.tensor("opt.world_model.step").item()) == 2
    finally:
        e.close()


def test_python_mirror_checks_what_the_reference_asserts():
    """dreamer_model.py / train_one_step.py assert ranks, the horizon and the model size table; the mirror raises at the
    same places instead of launching with a wrong shape."""
    from dreamer_v3_b200.models.components import MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from dreamer_v3_b200.utils.model_dimensions import get_gru_units
    with pytest.raises(AssertionError):
        get_gru_units("XXL")
    wm = WorldModel(model_dimension="nano", action_space=Discrete(2), batch_length_T=4, encoder=MLP(model_dimension="nano"),
                    decoder=VectorDecoder(model_dimension="nano", observation_space=Box(-1, 1, (3,))))
    with pytest.raises(RuntimeError, match="batch size unknown"):
        wm.engine
    dm = DreamerModel(model_dimension="nano", action_space=Discrete(2), batch_size_B=2, batch_length_T=4, horizon_H=2, world_model=wm)
    e = dm.engine
    try:
        with pytest.raises(RuntimeError, match="already built"):
            wm._configure(B=3)
        with pytest.raises(AssertionError):
            dm.dream_trajectory(start_states={"h": torch.zeros(8, e.cfg.G), "z": torch.zeros(8, e.cfg.Zc, e.cfg.Zk)},
                                start_is_terminated=torch.zeros(8), timesteps_H=5, gamma=0.997)
        with pytest.raises(ValueError, match="rows"):
            init = dm.get_initial_state()
            dm.forward_inference({k: v.repeat_interleave(3, 0) for k, v in init.items()}, np.zeros((3, 3), np.float32), np.ones(3, np.float32))
        with pytest.raises(RuntimeError):   # a sample of the wrong shape
            e.set_sample({"obs": np.zeros((2, 5, 3), np.float32)})
    finally:
        e.close()
