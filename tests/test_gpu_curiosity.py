"""GPU parity of the curiosity path: the Plan2Explore disagree nets (models/disagree_networks.py),
the intrinsic rewards inside the lambda-returns (critic_loss.py:118-120), disagree_loss (losses/disagree_loss.py) and the
fourth optimizer (train_one_step.py:201-246) against the oracle, which tests/test_reference_golden.py pins to the
reference's own code (fixtures ref_*_cur_*.npz; the CUDA engine is also held to those fixtures there).

fp32 mode: indices exact, tensors / losses / gradients 1e-4 relative (max-norm), as everywhere. Tensor-core mode: the
disagree quantities within the bf16 tolerance class of tests/test_gpu_configs.py (3e-2 of the tensor's largest value)."""
import numpy as np
import pytest
import torch

from dreamer_v3_b200.spec import EngineConfig

pytestmark = pytest.mark.gpu

TOL = 1e-4
DISAGREE_KEYS = {"DISAGREE_L_total", "DISAGREE_intrinsic_rewards_H_B", "DISAGREE_intrinsic_rewards", "DISAGREE_gradients_maxabs",
                 "DISAGREE_gradients_clipped_by_glob_norm_maxabs"}


@pytest.mark.parametrize("kind", ("xs-disc", "xs-box", "s-disc"))
def test_curiosity_update_parity_fp32(kind):
    """Two full updates with curiosity on at real layer widths (XS: D = G = 256, 1 layer; S: D = G = 512, 2 layers), reduced
    rows; 3 nets instead of 8 keep the fp64 oracle fast."""
    from _parity import ParityRun
    size, act = kind.split("-")
    cfg = EngineConfig.make(size.upper(), A=3 if act == "disc" else 2, discrete=act == "disc", obs_dim=5, B=4, T=8, H=3,
                            use_curiosity=True, num_curiosity_nets=3, intrinsic_rewards_scale=0.7)
    assert cfg.num_curiosity_nets == 3
    run = ParityRun(cfg, steps=2)
    try:
        assert all(v == 0 for v in run.exact.values()), run.exact
        for k in ("dream.z_predicted_probs", "dream.rewards_intrinsic", "ac.DISAGREE_L_total", "ac.value_targets",
                  "ac.disagree_grad_global_norm", "grad.disagree.net2.mlp0.w", "grad.disagree.net0.repr.bias"):
            assert k in run.report, k
        bad = {k: v for k, v in run.report.items() if not v <= TOL and not k.startswith("adam_delta.")}
        assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:20]
        bad = {k: v for k, v in run.report.items() if k.startswith("adam_delta.") and not v <= 1e-3}
        assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:20]
    finally:
        run.close()


def test_curiosity_changes_the_value_targets():
    """The intrinsic rewards must reach the lambda-returns (and only them: dream.rewards stays the extrinsic prediction)."""
    from _parity import ParityRun
    kw = dict(A=3, discrete=True, obs_dim=5, B=4, T=8, H=3)
    on = ParityRun(EngineConfig.make("XXS", use_curiosity=True, num_curiosity_nets=2, intrinsic_rewards_scale=5.0, **kw), steps=1)
    off = ParityRun(EngineConfig.make("XXS", **kw), steps=1)
    try:
        t_on, t_off = on.engine.tensor, off.engine.tensor
        # same weights, same noise: the same dream (up to the summation order of split-K atomics)
        assert torch.allclose(t_on("dream.rewards"), t_off("dream.rewards"), rtol=1e-4, atol=1e-5), \
            float((t_on("dream.rewards") - t_off("dream.rewards")).abs().max())
        ri = t_on("dream.rewards_intrinsic")
        assert float(ri.abs().max()) > 0 and abs(float(ri.mean())) < 1e-4 * float(ri.abs().max()) + 1e-6, \
            (float(ri.mean()), float(ri.abs().max()))   # centred over all rows
        assert float((t_on("ac.value_targets") - t_off("ac.value_targets")).abs().max()) > 1e-3 * float(ri.abs().max())
    finally:
        on.close(); off.close()


@pytest.mark.parametrize("discrete", (True, False))
def test_curiosity_api_matches_oracle_over_three_updates(discrete):
    """train_actor_and_critic_one_step(use_curiosity=True) through the reference-facing API: the one-graph path (dreaming
    from the engine's own observe pass) against the oracle over three updates, result keys of train_one_step.py:201-231."""
    from dreamer_v3_b200.models.components import MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from dreamer_v3_b200.training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step
    from oracle.dreamer_oracle import Oracle, make_synthetic_sample
    B, T, H = 4, 8, 4
    space = Discrete(3) if discrete else Box(-1, 1, (2,))
    wm = WorldModel(model_dimension="XXS", action_space=space, batch_length_T=T, encoder=MLP(model_dimension="XXS"),
                    decoder=VectorDecoder(model_dimension="XXS", observation_space=Box(-1, 1, (5,))), symlog_obs=True, seed=3)
    dm = DreamerModel(model_dimension="XXS", action_space=space, batch_size_B=B, batch_length_T=T, horizon_H=H, world_model=wm,
                      use_curiosity=True, intrinsic_rewards_scale=0.3)
    e = dm.engine
    cfg = e.cfg
    assert cfg.num_curiosity_nets == 8 and dm.disagree_nets.num_networks == 8
    assert len(dm.disagree_nets.trainable_variables) == 8 * (3 * cfg.L + 2)
    g = torch.Generator().manual_seed(0)
    for n in ("rew.head.w", "critic.head.w"):
        e.tensor(n).copy_(0.02 * torch.randn(e.tensor(n).shape, generator=g))
    dm.critic.init_ema()
    oracle = Oracle(cfg, e.get_params(), dtype=torch.float64)
    sample = make_synthetic_sample(cfg, seed=5)
    rng = np.random.default_rng(6)
    with pytest.raises(ValueError):
        train_actor_and_critic_one_step(
            world_model_forward_train_outs={"h_states_BxT": None}, is_terminated=None, horizon_H=H, gamma=0.997, lambda_=0.95,
            actor_grad_clip=100.0, critic_grad_clip=100.0, disagree_grad_clip=100.0, dreamer_model=dm, entropy_scale=3e-4,
            return_normalization_decay=0.99, use_curiosity=False)
    for step in range(3):
        ores = oracle.train_world_model_one_step(sample, rng=rng, margin=1e-3)
        wm.set_noise(u_post=oracle.last_noise["u_post"])
        res = train_world_model_one_step(sample=sample, batch_size_B=B, batch_length_T=T, grad_clip=1000.0, world_model=wm)
        fwd, ofwd = res["WORLD_MODEL_forward_train_outs"], ores["WORLD_MODEL_forward_train_outs"]
        oac = oracle.train_actor_and_critic_one_step(ofwd["h_states_BxT"], ofwd["z_states_BxT"], sample["is_terminated"], H=H,
                                                     rng=rng, margin=1e-3)
        wm.set_noise(u_dyn=oracle.last_noise["u_dyn"], act_noise=oracle.last_noise["u_act" if discrete else "eps_act"])
        ac = train_actor_and_critic_one_step(
            world_model_forward_train_outs=fwd, is_terminated=torch.as_tensor(sample["is_terminated"]).reshape(-1),
            horizon_H=H, gamma=0.997, lambda_=0.95, actor_grad_clip=100.0, critic_grad_clip=100.0, disagree_grad_clip=100.0,
            dreamer_model=dm, entropy_scale=3e-4, return_normalization_decay=0.99, use_curiosity=True)
        assert DISAGREE_KEYS <= set(ac)
        d, od = ac["dream_data"], oac["dream_data"]
        assert {"rewards_intrinsic_t1_to_H_B", "z_predicted_probs_N_HxB"} <= set(d)
        assert len(d["z_predicted_probs_N_HxB"]) == 8 and tuple(d["z_predicted_probs_N_HxB"][0].shape) == ((H + 1) * cfg.N, cfg.Zc, cfg.Zk)
        ri, ori = d["rewards_intrinsic_t1_to_H_B"].cpu().numpy(), od["rewards_intrinsic_t1_to_H_B"].detach().numpy()
        assert ri.shape == (H, cfg.N)
        np.testing.assert_allclose(ri, ori, atol=2e-4 * np.abs(ori).max())
        np.testing.assert_allclose(float(ac["DISAGREE_L_total"]), float(oac["DISAGREE_L_total"]), rtol=2e-4)
        np.testing.assert_allclose(float(ac["DISAGREE_intrinsic_rewards"]), float(oac["DISAGREE_intrinsic_rewards"]),
                                   atol=2e-4 * np.abs(ori).max())
        np.testing.assert_allclose(float(ac["DISAGREE_gradients_maxabs"]), oac["DISAGREE_gradients_maxabs"], rtol=1e-3)
        np.testing.assert_allclose(float(ac["DISAGREE_gradients_clipped_by_glob_norm_maxabs"]),
                                   oac["DISAGREE_gradients_clipped_by_glob_norm_maxabs"], rtol=1e-3)
        np.testing.assert_allclose(ac["VALUE_TARGETS_H_B"].cpu().numpy(), oac["VALUE_TARGETS_H_B"].detach().numpy(), rtol=1e-4, atol=1e-5)
        for k in ("CRITIC_L_total", "ACTOR_L_total"):
            np.testing.assert_allclose(float(ac[k]), float(oac[k]), rtol=2e-4, atol=1e-6)
    assert dm.disagree_nets.optimizer.iterations == 3
    # the weights after three Adam steps (lr 1e-4, eps 1e-5): the update itself against the oracle's
    op = oracle.numpy_params()
    for n in ("disagree.net0.mlp0.w", "disagree.net7.repr.w", "disagree.net3.mlp0.g"):
        np.testing.assert_allclose(e.tensor(n).cpu().numpy(), op[n], atol=2e-6)
    e.close()


def test_curiosity_tensor_core_mode():
    """compute_mode 1 (bf16 tcgen05 contractions) with >= 128 rows so that the disagree nets take the tensor-core kernels:
    the device's draws are forced on the oracle (see ParityRun), the disagree tensors within the bf16 tolerance class."""
    from _parity import ParityRun
    cfg = EngineConfig.make("XXS", A=3, discrete=True, obs_dim=5, B=8, T=16, H=3, use_curiosity=True, num_curiosity_nets=3,
                            intrinsic_rewards_scale=0.7)
    run = ParityRun(cfg, steps=1, compute_mode=1, forced=True)
    try:
        rep = run.report
        print("curiosity bf16:", {k: round(v, 5) for k, v in rep.items() if "isagree" in k or "intrinsic" in k or "z_predicted" in k})
        for k in ("dream.z_predicted_probs", "ac.DISAGREE_L_total"):   # forward class of tests/test_gpu_configs.py
            assert rep[k] <= 2e-2, (k, rep[k])
        # a centred statistic of small differences between nets: bounded against its own largest value, looser
        assert rep["dream.rewards_intrinsic"] <= 1e-1, rep["dream.rewards_intrinsic"]
        bad = {k: v for k, v in rep.items() if (k.startswith("grad.disagree.") or k == "ac.disagree_grad_global_norm") and not v <= 6e-2}
        assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:10]
        assert max(v for k, v in rep.items() if k.startswith("grad.disagree.")) > 1e-5   # the tensor-core path really ran
    finally:
        run.close()


def test_curiosity_checkpoint_round_trip(tmp_path):
    """run_experiment.py:692-714 / :489-509 with use_curiosity: disagree_nets.npz + disagree_nets_optimizer.npz."""
    from dreamer_v3_b200.models.components import MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from dreamer_v3_b200.utils import checkpoint

    def build(seed):
        wm = WorldModel(model_dimension="nano", action_space=Discrete(2), batch_length_T=4, encoder=MLP(model_dimension="nano"),
                        decoder=VectorDecoder(model_dimension="nano", observation_space=Box(-1, 1, (3,))), seed=seed)
        return DreamerModel(model_dimension="nano", action_space=Discrete(2), batch_size_B=2, batch_length_T=4, horizon_H=2,
                            world_model=wm, use_curiosity=True)
    a, b = build(1), build(2)
    ea, eb = a.engine, b.engine
    ea.tensor("opt.disagree.step").fill_(7)
    ea.tensor("adam_m.disagree.net1.repr.w").normal_()
    import os
    for exact in (True, False):
        path = str(tmp_path / ("exact" if exact else "keras"))
        checkpoint.save_checkpoint(path, a, iteration=3)
        assert os.path.exists(os.path.join(path, "disagree_nets.npz")) and os.path.exists(os.path.join(path, "disagree_nets_optimizer.npz"))
        if not exact:
            os.remove(os.path.join(path, "dv3_state.npz"))
        assert checkpoint.load_checkpoint(path, b)["iteration"] == 3
        for n in ("disagree.net0.mlp0.w", "disagree.net7.repr.bias", "adam_m.disagree.net1.repr.w"):
            assert torch.equal(ea.tensor(n), eb.tensor(n)), n
        assert int(eb.tensor("opt.disagree.step").item()) == 7
        eb.tensor("disagree.net0.mlp0.w").zero_()
    ea.close(); eb.close()
