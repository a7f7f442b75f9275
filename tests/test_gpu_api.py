"""The reference-facing Python API (WorldModel / DreamerModel / train_*_one_step) on the GPU: result-dict keys of
training/train_one_step.py:88-126, critic_loss.py:78-89, actor_loss.py:81-96 and values against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WM_KEYS = {
    "WORLD_MODEL_forward_train_outs", "WORLD_MODEL_learned_initial_h", "WORLD_MODEL_L_decoder_B_T", "WORLD_MODEL_L_decoder",
    "WORLD_MODEL_L_reward_B_T", "WORLD_MODEL_L_reward", "WORLD_MODEL_L_continue_B_T", "WORLD_MODEL_L_continue",
    "WORLD_MODEL_L_prediction_B_T", "WORLD_MODEL_L_prediction", "WORLD_MODEL_L_dynamics_B_T", "WORLD_MODEL_L_dynamics",
    "WORLD_MODEL_L_representation_B_T", "WORLD_MODEL_L_representation", "WORLD_MODEL_L_total_B_T", "WORLD_MODEL_L_total",
    "WORLD_MODEL_gradients_maxabs", "WORLD_MODEL_gradients_clipped_by_glob_norm_maxabs"}
FWD_KEYS = {"sampled_obs_symlog_BxT", "obs_distribution_BxT", "reward_logits_BxT", "rewards_BxT", "continue_distribution_BxT",
            "continues_BxT", "z_posterior_probs_BxT", "z_prior_probs_BxT", "h_states_BxT", "z_states_BxT"}
AC_KEYS = {
    "VALUE_TARGETS_symlog_H_B", "CRITIC_L_total", "CRITIC_L_total_H_B", "CRITIC_L_neg_logp_of_value_targets_H_B",
    "CRITIC_L_neg_logp_of_value_targets", "CRITIC_L_slow_critic_regularization_H_B", "CRITIC_L_slow_critic_regularization",
    "VALUE_TARGETS_H_B", "dream_data", "ACTOR_L_total_H_B", "ACTOR_L_total", "ACTOR_logp_actions_dreamed_H_B",
    "ACTOR_scaled_value_targets_H_B", "ACTOR_value_targets_pct95_ema", "ACTOR_value_targets_pct5_ema",
    "ACTOR_action_entropy_H_B", "ACTOR_action_entropy", "ACTOR_L_neglogp_reinforce_term_H_B", "ACTOR_L_neglogp_reinforce_term",
    "ACTOR_L_neg_entropy_term_H_B", "ACTOR_L_neg_entropy_term", "ACTOR_gradients_maxabs",
    "ACTOR_gradients_clipped_by_glob_norm_maxabs", "CRITIC_gradients_maxabs", "CRITIC_gradients_clipped_by_glob_norm_maxabs"}
DREAM_KEYS = {"h_states_t0_to_H_B", "z_states_prior_t0_to_H_B", "rewards_dreamed_t0_to_H_B", "continues_dreamed_t0_to_H_B",
              "actions_dreamed_t0_to_H_B", "actions_dreamed_distributions_t0_to_H_B", "values_dreamed_t0_to_H_B",
              "values_symlog_dreamed_logits_t0_to_HxB", "v_symlog_dreamed_ema_t0_to_H_B", "dream_loss_weights_t0_to_H_B"}


@pytest.mark.parametrize("discrete", [True, False])
def test_api_matches_oracle_over_three_updates(discrete):
    from dreamer_v3_b200.models.components import MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from dreamer_v3_b200.training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step
    from oracle.dreamer_oracle import Oracle, make_synthetic_sample
    B, T, H = 4, 8, 5
    space = Discrete(3) if discrete else Box(-1, 1, (2,))
    wm = WorldModel(model_dimension="XXS", action_space=space, batch_length_T=T, encoder=MLP(model_dimension="XXS"),
                    decoder=VectorDecoder(model_dimension="XXS", observation_space=Box(-1, 1, (5,))), symlog_obs=True, seed=3)
    dm = DreamerModel(model_dimension="XXS", action_space=space, batch_size_B=B, batch_length_T=T, horizon_H=H, world_model=wm)
    e = dm.engine
    cfg = e.cfg
    # break the symmetry of the zero-initialised heads so that every loss term is exercised
    g = torch.Generator().manual_seed(0)
    for n in ("rew.head.w", "critic.head.w"):
        e.tensor(n).copy_(0.02 * torch.randn(e.tensor(n).shape, generator=g))
    dm.critic.init_ema()
    oracle = Oracle(cfg, e.get_params(), dtype=torch.float64)
    sample = make_synthetic_sample(cfg, seed=5)
    rng = np.random.default_rng(6)
    assert set(dm.get_initial_state()) == {"h", "z", "a"}
    for step in range(3):
        ores = oracle.train_world_model_one_step(sample, rng=rng, margin=1e-3)
        wm.set_noise(u_post=oracle.last_noise["u_post"])
        res = train_world_model_one_step(sample=sample, batch_size_B=B, batch_length_T=T, grad_clip=1000.0, world_model=wm)
        assert set(res) == WM_KEYS and set(res["WORLD_MODEL_forward_train_outs"]) == FWD_KEYS
        for k in ("WORLD_MODEL_L_total", "WORLD_MODEL_L_decoder", "WORLD_MODEL_L_reward", "WORLD_MODEL_L_continue",
                  "WORLD_MODEL_L_dynamics", "WORLD_MODEL_L_prediction"):
            np.testing.assert_allclose(float(res[k]), float(ores[k]), rtol=1e-4)
        np.testing.assert_allclose(res["WORLD_MODEL_L_total_B_T"].cpu().numpy(), ores["WORLD_MODEL_L_total_B_T"].detach().numpy(), rtol=1e-4)
        fwd = res["WORLD_MODEL_forward_train_outs"]
        ofwd = ores["WORLD_MODEL_forward_train_outs"]
        np.testing.assert_allclose(fwd["h_states_BxT"].cpu().numpy(), ofwd["h_states_BxT"].numpy(), atol=1e-4)
        np.testing.assert_array_equal(fwd["z_states_BxT"].argmax(-1).cpu().numpy(), ofwd["z_indices_BxT"])
        np.testing.assert_allclose(fwd["obs_distribution_BxT"].loc.cpu().numpy(), ofwd["obs_distribution_BxT.loc"].numpy(), atol=1e-4)
        np.testing.assert_allclose(float(res["WORLD_MODEL_gradients_maxabs"]), ores["WORLD_MODEL_gradients_maxabs"], rtol=1e-3)
        oac = oracle.train_actor_and_critic_one_step(ofwd["h_states_BxT"], ofwd["z_states_BxT"], sample["is_terminated"], H=H,
                                                     rng=rng, margin=1e-3)
        wm.set_noise(u_dyn=oracle.last_noise["u_dyn"], act_noise=oracle.last_noise["u_act" if discrete else "eps_act"])
        ac = train_actor_and_critic_one_step(
            world_model_forward_train_outs=fwd, is_terminated=torch.as_tensor(sample["is_terminated"]).reshape(-1),
            horizon_H=H, gamma=0.997, lambda_=0.95, actor_grad_clip=100.0, critic_grad_clip=100.0, disagree_grad_clip=100.0,
            dreamer_model=dm, entropy_scale=3e-4, return_normalization_decay=0.99)
        assert set(ac) == AC_KEYS
        assert DREAM_KEYS <= set(ac["dream_data"])
        for k in ("CRITIC_L_total", "ACTOR_L_total", "ACTOR_action_entropy", "CRITIC_L_neg_logp_of_value_targets"):
            np.testing.assert_allclose(float(ac[k]), float(oac[k]), rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(ac["VALUE_TARGETS_H_B"].cpu().numpy(), oac["VALUE_TARGETS_H_B"].detach().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(float(ac["ACTOR_value_targets_pct95_ema"]), oac["ACTOR_value_targets_pct95_ema"], rtol=1e-4, atol=1e-6)
        d, od = ac["dream_data"], oac["dream_data"]
        np.testing.assert_allclose(d["rewards_dreamed_t0_to_H_B"].cpu().numpy(), od["rewards_dreamed_t0_to_H_B"].detach().numpy(), atol=1e-4)
        np.testing.assert_allclose(d["dream_loss_weights_t0_to_H_B"].cpu().numpy(), od["dream_loss_weights_t0_to_H_B"].detach().numpy(), atol=1e-5)
        dist0 = d["actions_dreamed_distributions_t0_to_H_B"][0]
        assert dist0.entropy().shape == (cfg.N,)
    assert dm.critic.optimizer.iterations == 3 and wm.optimizer.iterations == 3
    e.close()


@pytest.mark.parametrize("kind", ["vec-disc", "vec-box", "img-disc"])
def test_forward_inference_matches_oracle(kind):
    """Acting step: DreamerModel.forward_inference over a short closed loop, with is_first resets."""
    from dreamer_v3_b200.models.components import CNNAtari, ConvTransposeAtari, MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from oracle.dreamer_oracle import Oracle
    image = kind.startswith("img")
    discrete = kind.endswith("disc")
    size = "micro" if image else "XXS"
    space = Discrete(4) if discrete else Box(-1, 1, (3,))
    if image:
        enc, dec = CNNAtari(model_dimension=size), ConvTransposeAtari(model_dimension=size, gray_scaled=False)
    else:
        enc, dec = MLP(model_dimension=size), VectorDecoder(model_dimension=size, observation_space=Box(-1, 1, (5,)))
    wm = WorldModel(model_dimension=size, action_space=space, batch_length_T=4, encoder=enc, decoder=dec,
                    symlog_obs=not image, seed=1)
    dm = DreamerModel(model_dimension=size, action_space=space, batch_size_B=4, batch_length_T=4, horizon_H=3, world_model=wm)
    e = dm.engine
    cfg = e.cfg
    g = torch.Generator().manual_seed(0)
    e.tensor("initial_h").copy_(0.3 * torch.randn(cfg.G, generator=g))   # non-trivial learned initial state
    oracle = Oracle(cfg, e.get_params(), dtype=torch.float64)
    rng = np.random.default_rng(1)
    rows = 3
    init = dm.get_initial_state()
    assert init["h"].shape == (1, cfg.G) and init["z"].shape == (1, cfg.Zc, cfg.Zk) and init["a"].shape == (1, cfg.A)
    oinit = oracle.get_initial_state()
    np.testing.assert_allclose(init["h"].cpu().numpy(), oinit["h"].detach().numpy(), atol=1e-6)
    np.testing.assert_array_equal(init["z"].cpu().numpy(), oinit["z"].numpy())
    states = {k: v.repeat(rows, *([1] * (v.dim() - 1))) for k, v in init.items()}
    ostates = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in states.items()}
    for step in range(5):
        obs = rng.standard_normal((rows,) + ((64, 64, 3) if image else (5,))).astype(np.float32)
        is_first = np.array([1.0 if step == 0 else 0.0, 1.0 if step == 3 else 0.0, 0.0], np.float32)
        o = oracle.forward_inference(ostates["h"], ostates["z"], ostates["a"], obs, is_first, rng=rng, margin=1e-3)
        wm.set_inference_noise(oracle.last_noise.pop("inf_u_z")[0], oracle.last_noise.pop("inf_act_noise")[0])
        actions, new = dm.forward_inference(states, obs, is_first)
        np.testing.assert_allclose(new["h"].cpu().numpy(), o["h"].numpy(), atol=1e-4)
        np.testing.assert_array_equal(new["z"].argmax(-1).cpu().numpy(), o["z_idx"])
        np.testing.assert_allclose(actions.cpu().numpy(), o["a"].numpy(), atol=1e-4)
        assert set(new) == {"h", "z", "a"}
        states = new
        ostates = {"h": o["h"].numpy(), "z": o["z"].numpy(), "a": o["a"].numpy()}
    wz = wm.forward_inference(states, obs, is_first)
    assert set(wz) == {"h", "z"}
    e.close()


@pytest.mark.gpu
def test_checkpoint_round_trip_in_the_reference_layout(tmp_path):
    """run_experiment.py:692-714 / :489-509 file layout (`weights=model.get_weights()` object arrays +
    optimizer lists + counters) and the exact name-keyed state; two updates, save, two more, restore, two more -> the
    same parameters."""
    import os
    from dreamer_v3_b200.models.components import MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box
    from dreamer_v3_b200.training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step
    from dreamer_v3_b200.utils.checkpoint import load_checkpoint, save_checkpoint
    from oracle.dreamer_oracle import make_synthetic_sample
    B, T, H = 4, 8, 5
    space = Box(-1, 1, (2,))
    wm = WorldModel(model_dimension="XXS", action_space=space, batch_length_T=T, encoder=MLP(model_dimension="XXS"),
                    decoder=VectorDecoder(model_dimension="XXS", observation_space=Box(-1, 1, (5,))), symlog_obs=True, seed=5)
    dm = DreamerModel(model_dimension="XXS", action_space=space, batch_size_B=B, batch_length_T=T, horizon_H=H, world_model=wm)
    e = dm.engine
    sample = make_synthetic_sample(e.cfg, seed=9)
    hp = dict(horizon_H=H, gamma=0.997, lambda_=0.95, actor_grad_clip=100.0, critic_grad_clip=100.0, disagree_grad_clip=100.0,
              entropy_scale=3e-4, return_normalization_decay=0.99)

    def updates(n):
        for _ in range(n):
            r = train_world_model_one_step(sample=sample, batch_size_B=B, batch_length_T=T, grad_clip=1000.0, world_model=wm)
            train_actor_and_critic_one_step(world_model_forward_train_outs=r["WORLD_MODEL_forward_train_outs"],
                                            is_terminated=sample["is_terminated"].reshape(-1), dreamer_model=dm, **hp)
    rng = np.random.default_rng(4)
    cfg = e.cfg
    wm.set_noise(u_post=rng.random((T, B, cfg.Zc), dtype=np.float32), u_dyn=rng.random((H, cfg.N, cfg.Zc), dtype=np.float32),
                 act_noise=rng.standard_normal((H + 1, cfg.N, cfg.A)).astype(np.float32))  # same noise every update
    try:
        updates(2)
        save_checkpoint(str(tmp_path), dm, total_env_steps=123, total_replayed_steps=456, total_train_steps=2, iteration=1)
        assert len(wm.get_weights()) == len(list(wm.trainable_variables)) and wm.optimizer.get_weights()[0] == 2
        updates(2)
        want = e.get_params()
        for route in ("exact", "reference layout"):
            if route == "reference layout":
                os.remove(os.path.join(str(tmp_path), "dv3_state.npz"))
            for n in e.tensor_names():  # scramble everything a checkpoint has to restore
                if n.startswith(("adam_m.", "adam_v.")) or n in want:
                    e.tensor(n).normal_()
            counters = load_checkpoint(str(tmp_path), dm)
            assert counters == dict(total_env_steps=123, total_replayed_steps=456, total_train_steps=2, iteration=1)
            updates(2)
            got = e.get_params()
            for k in want:  # (gradient accumulation uses fp32 atomics: equal up to summation order, not bit for bit)
                np.testing.assert_allclose(got[k], want[k], rtol=1e-4, atol=1e-6, err_msg=f"{route}: {k}")
    finally:
        e.close()
