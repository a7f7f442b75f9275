"""Golden vectors produced by the REFERENCE'S OWN CODE.

TensorFlow / Keras / tfp cannot be installed in the build image, so the reference is executed over the API shim in
`tests/golden/tfshim/` (PyTorch-CPU backed `tensorflow`, `tensorflow_probability`, `gymnasium`, `tree`): the files
under /root/reference - models/world_model.py, models/dreamer_model.py, models/components/*, models/actor_network.py,
models/critic_network.py, losses/*, training/train_one_step.py, utils/symlog.py, utils/two_hot.py - are imported
UNMODIFIED and `train_world_model_one_step` / `train_actor_and_critic_one_step` are called exactly as
run_experiment.py:426-472 calls them.  What is pinned: the reference's composition (masking, wrap-around action
index, concatenation orders, stop-gradients, loss weighting, lambda-returns, percentile EMA, clipping, optimizer and
EMA calls).  What is NOT pinned: TensorFlow's own arithmetic of the primitives, which the shim restates (SURVEY.md
Appendix A.1) - see DESIGN.md section 5.

    python tests/golden/make_reference_golden.py [case ...]   # needs /root/reference; writes tests/golden/ref_*.npz

Each fixture holds: parameters before the update (spec names), the replay sample, every random number the
reference consumed (uniforms of the categorical draws, normals of the Box actor), and the reference's outputs:
forward_train tensors, losses, dream tensors, value targets, actor / critic losses, all gradients, and the parameters
after one and after two updates (Adam t = 1, 2; critic EMA; percentile EMA).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DV3_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tfshim"))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import gymnasium as gym  # noqa: E402  (shim)
import tensorflow as tf  # noqa: E402  (shim)
import tensorflow_probability as tfp  # noqa: E402  (shim)

from dreamer_v3_b200.spec import EngineConfig, init_params, param_spec  # noqa: E402  (names / shapes only)

CASES = {
    # name: (engine config kwargs, seed)
    "vec_disc": (dict(model_dimension="nano", A=2, discrete=True, obs_dim=4, B=3, T=5, H=4), 101),
    "vec_box": (dict(model_dimension="nano", A=2, discrete=False, obs_dim=3, B=3, T=5, H=4), 202),
    "img_disc": (dict(model_dimension="nano", A=3, discrete=True, image_obs=True, B=2, T=3, H=2), 303),
    # use_curiosity=True: the Plan2Explore disagree nets (models/disagree_networks.py, losses/disagree_loss.py), 8 nets at nano
    "vec_disc_cur": (dict(model_dimension="nano", A=3, discrete=True, obs_dim=4, B=3, T=4, H=3, use_curiosity=True,
                          intrinsic_rewards_scale=0.1), 404),
    "vec_box_cur": (dict(model_dimension="nano", A=2, discrete=False, obs_dim=3, B=2, T=4, H=3, use_curiosity=True,
                         intrinsic_rewards_scale=0.5), 505),
}


def build_reference(cfg):
    """Construction as in run_experiment.py:185-212."""
    from models.components.cnn_atari import CNNAtari
    from models.components.conv_transpose_atari import ConvTransposeAtari
    from models.components.mlp import MLP
    from models.components.vector_decoder import VectorDecoder
    from models.dreamer_model import DreamerModel
    from models.world_model import WorldModel

    md = cfg.model_dimension
    action_space = gym.spaces.Discrete(cfg.A) if cfg.discrete else gym.spaces.Box(-1.0, 1.0, (cfg.A,))
    if cfg.image_obs:
        encoder = CNNAtari(model_dimension=md)
        decoder = ConvTransposeAtari(model_dimension=md, gray_scaled=False)
    else:
        encoder = MLP(model_dimension=md)
        decoder = VectorDecoder(model_dimension=md, observation_space=gym.spaces.Box(-10.0, 10.0, (cfg.obs_dim,)))
    world_model = WorldModel(model_dimension=md, action_space=action_space, batch_length_T=cfg.T, encoder=encoder,
                             decoder=decoder, symlog_obs=cfg.symlog_obs)
    return DreamerModel(model_dimension=md, action_space=action_space, world_model=world_model, batch_size_B=cfg.B,
                        batch_length_T=cfg.T, horizon_H=cfg.H, use_curiosity=cfg.num_curiosity_nets > 0,
                        intrinsic_rewards_scale=cfg.intrinsic_rewards_scale)


def variable_map(dm, cfg):
    """spec name (dreamer_v3_b200/spec.py) -> the reference model's tf.Variable, by attribute path."""
    wm, m = dm.world_model, {}

    def mlp(prefix, net, L):
        for l in range(L):
            m[f"{prefix}.mlp{l}.w"] = net.dense_layers[l].kernel
            m[f"{prefix}.mlp{l}.g"] = net.layer_normalizations[l].gamma
            m[f"{prefix}.mlp{l}.b"] = net.layer_normalizations[l].beta

    def dense(prefix, layer):
        m[f"{prefix}.w"], m[f"{prefix}.bias"] = layer.kernel, layer.bias

    if cfg.image_obs:
        for l in range(4):
            m[f"enc.conv{l}.w"] = wm.encoder.conv_layers[l].kernel
            m[f"enc.conv{l}.g"] = wm.encoder.layer_normalizations[l].gamma
            m[f"enc.conv{l}.b"] = wm.encoder.layer_normalizations[l].beta
        dense("dec.dense", wm.decoder.dense_layer)
        for l in range(3):
            m[f"dec.convT{l}.w"] = wm.decoder.conv_transpose_layers[l].kernel
            m[f"dec.convT{l}.g"] = wm.decoder.layer_normalizations[l].gamma
            m[f"dec.convT{l}.b"] = wm.decoder.layer_normalizations[l].beta
        dense("dec.out", wm.decoder.output_conv2d_transpose)
    else:
        mlp("enc", wm.encoder, cfg.L)
        mlp("dec", wm.decoder.mlp, cfg.L)
        dense("dec.out", wm.decoder.mlp.output_layer)
    mlp("post", wm.posterior_mlp, 1)
    dense("post.repr", wm.posterior_representation_layer.z_generating_layer)
    mlp("dyn", wm.dynamics_predictor.mlp, 1)
    dense("dyn.repr", wm.dynamics_predictor.representation_layer.z_generating_layer)
    m["initial_h"] = wm.initial_h
    mlp("seq.pre", wm.sequence_model.pre_gru_layer, 1)
    m["seq.gru.kernel"] = wm.sequence_model.gru_unit.kernel
    m["seq.gru.recurrent"] = wm.sequence_model.gru_unit.recurrent_kernel
    m["seq.gru.bias"] = wm.sequence_model.gru_unit.bias
    mlp("rew", wm.reward_predictor.mlp, cfg.L)
    dense("rew.head", wm.reward_predictor.reward_layer.reward_buckets_layer)
    mlp("cont", wm.continue_predictor.mlp, cfg.L)
    dense("cont.out", wm.continue_predictor.mlp.output_layer)
    mlp("actor", dm.actor.mlp, cfg.L)
    dense("actor.out", dm.actor.mlp.output_layer)
    if not cfg.discrete:
        mlp("actor.std", dm.actor.std_mlp, cfg.L)
        dense("actor.std_out", dm.actor.std_mlp.output_layer)
    mlp("critic", dm.critic.mlp, cfg.L)
    dense("critic.head", dm.critic.return_layer.reward_buckets_layer)
    mlp("critic_ema", dm.critic.mlp_ema, cfg.L)
    dense("critic_ema.head", dm.critic.return_layer_ema.reward_buckets_layer)
    for n in range(cfg.num_curiosity_nets):
        mlp(f"disagree.net{n}", dm.disagree_nets.mlps[n], cfg.L)
        dense(f"disagree.net{n}.repr", dm.disagree_nets.representation_layers[n].z_generating_layer)
    return m


def make_sample(cfg, rng):
    B, T, A = cfg.B, cfg.T, cfg.A
    if cfg.image_obs:
        obs = rng.integers(0, 256, size=(B, T, 64, 64, 3), dtype=np.uint8).astype(np.float32) / 128.0 - 1.0
    else:
        obs = rng.standard_normal((B, T, cfg.obs_dim)) * 3.0
    if cfg.discrete:
        actions = np.eye(A)[rng.integers(0, A, size=(B, T))]
    else:
        actions = rng.uniform(-1.0, 1.0, size=(B, T, A))
    is_first = (rng.random((B, T)) < 0.2).astype(np.float32)
    is_first[:, 0] = 1.0
    is_first[0, 0] = 0.0  # a row that continues an episode: exercises the t=0 wrap-around of actions[t-1]
    is_terminated = np.zeros((B, T))
    is_terminated[:, :-1] = is_first[:, 1:]
    is_terminated[-1, -1] = 1.0
    rewards = rng.choice([0.0, 1.0, -1.0, 0.37, 12.5, -300.0], size=(B, T))
    return {k: v.astype(np.float32) for k, v in
            dict(obs=obs, actions=actions, rewards=rewards, is_first=is_first, is_terminated=is_terminated).items()}


def run_case(name, kw, seed, dtype):
    from training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step

    tf.set_float(dtype)
    cfg = EngineConfig.make(**kw)
    N, H, Zc = cfg.N, cfg.H, cfg.Zc
    rng = np.random.default_rng(seed)
    sample_np = make_sample(cfg, rng)
    sample = {k: tf.convert_to_tensor(v, dtype=dtype) for k, v in sample_np.items()}
    dm = build_reference(cfg)
    wm = dm.world_model
    cur = cfg.num_curiosity_nets > 0
    models = [("world_model", wm), ("actor", dm.actor), ("critic", dm.critic)] + ([("disagree", dm.disagree_nets)] if cur else [])
    grads_seen = []
    real_clip = tf.clip_by_global_norm

    def recording_clip(t_list, clip_norm):
        t_list = list(t_list)
        grads_seen.append(t_list)
        return real_clip(t_list, clip_norm)

    def one_update():
        """run_experiment.py:426-472 (hyper-parameters :131-148, :245-252)."""
        grads_seen.clear()
        tfp.DRAW_LOG.clear()
        r_wm = train_world_model_one_step(
            sample=sample, batch_size_B=tf.convert_to_tensor(cfg.B), batch_length_T=tf.convert_to_tensor(cfg.T),
            grad_clip=tf.convert_to_tensor(1000.0), world_model=wm)
        fwd = r_wm["WORLD_MODEL_forward_train_outs"]
        n_wm_draws = len(tfp.DRAW_LOG)
        r_ac = train_actor_and_critic_one_step(
            world_model_forward_train_outs=fwd, is_terminated=tf.reshape(sample["is_terminated"], [-1]),
            horizon_H=H, gamma=0.997, lambda_=0.95, actor_grad_clip=100.0, critic_grad_clip=100.0,
            disagree_grad_clip=100.0, dreamer_model=dm, entropy_scale=3e-4, return_normalization_decay=0.99,
            train_actor=True, use_curiosity=cur)
        return r_wm, r_ac, n_wm_draws

    # ---- build pass: creates the lazily-built Keras variables; its numbers are discarded ----
    tfp.RNG = np.random.default_rng(seed + 1)
    tf.clip_by_global_norm = recording_clip
    try:
        one_update()
        vm = variable_map(dm, cfg)
        spec = {p.name: p for p in param_spec(cfg)}
        assert set(vm) == set(spec), sorted(set(vm) ^ set(spec))
        params0 = init_params(cfg, seed=seed + 2, mode="random")
        for n, var in vm.items():
            assert tuple(var.shape) == tuple(spec[n].shape), (n, tuple(var.shape), spec[n].shape)
            var.assign(torch.as_tensor(params0[n], dtype=dtype))
        # every trainable variable of the reference is in the table, in the right optimizer group
        by_id = {id(v): n for n, v in vm.items()}
        for grp, model in models:
            got = sorted(by_id[id(v)] for v in model.trainable_variables)
            want = sorted(p.name for p in param_spec(cfg) if p.group == grp)
            assert got == want, (grp, sorted(set(got) ^ set(want)))
        for opt in [m.optimizer for _, m in models]:  # forget the build pass
            opt.iterations, opt.m, opt.v = 0, {}, {}
        dm.actor.ema_value_target_pct5.assign(np.nan)
        dm.actor.ema_value_target_pct95.assign(np.nan)

        out = {"meta/case": np.array(name), "meta/dtype": np.array(str(dtype).replace("torch.", "")),
               "meta/cfg": np.array(repr(sorted(kw.items())))}
        out.update({"param0/" + n: params0[n] for n in vm})
        out.update({"sample/" + k: v for k, v in sample_np.items()})
        tfp.RNG = np.random.default_rng(seed + 3)
        npf = lambda x: x.numpy().astype(np.float64 if dtype == torch.float64 else np.float32)  # noqa: E731
        for step in (1, 2):
            r_wm, r_ac, n_wm = one_update()
            pre = f"s{step}/"
            # ---- noise in the layouts of include/dv3.h --------------------------------------------
            log = list(tfp.DRAW_LOG)
            wm_log, ac_log = log[:n_wm], log[n_wm:]
            # forward_train: initial-state draw, then per t (posterior, prior), then the decoder's obs sample
            assert len(wm_log) == 1 + 2 * cfg.T + 1 and wm_log[0][1].shape == (1, Zc)
            out[pre + "noise/u_post"] = np.stack([wm_log[1 + 2 * t][1] for t in range(cfg.T)], 0)      # [T,B,Zc]
            # dream_trajectory: actor(t0), then per step (prior z, actor)
            # (+ with curiosity one draw per disagree net after the rollout: their RepresentationLayers sample a z that
            # compute_intrinsic_rewards discards, disagree_networks.py:88-91 - not part of the noise contract)
            assert len(ac_log) == 1 + 2 * H + cfg.num_curiosity_nets
            assert all(d[1].shape == ((H + 1) * N, Zc) for d in ac_log[1 + 2 * H:])
            out[pre + "noise/act"] = np.stack([ac_log[0][1]] + [ac_log[2 + 2 * i][1] for i in range(H)], 0)
            out[pre + "noise/u_dyn"] = np.stack([ac_log[1 + 2 * i][1] for i in range(H)], 0)          # [H,N,Zc]
            assert out[pre + "noise/u_dyn"].shape == (H, N, Zc)
            # ---- world model ----------------------------------------------------------------------
            fwd = r_wm["WORLD_MODEL_forward_train_outs"]
            out[pre + "wm/obs_loc"] = npf(fwd["obs_distribution_BxT"].loc)
            out[pre + "wm/continue_logits"] = npf(fwd["continue_distribution_BxT"].logits)
            for k in ("reward_logits_BxT", "rewards_BxT", "continues_BxT", "z_posterior_probs_BxT",
                      "z_prior_probs_BxT", "h_states_BxT"):
                out[pre + "wm/" + k] = npf(fwd[k])
            out[pre + "wm/z_indices_BxT"] = fwd["z_states_BxT"].numpy().argmax(-1).astype(np.int32)
            for k, v in r_wm.items():
                if k.startswith("WORLD_MODEL_L_") or k.endswith("_maxabs"):
                    out[pre + "wm/" + k] = npf(v)
            # ---- dream / actor / critic ------------------------------------------------------------
            dd = r_ac["dream_data"]
            for k in ("h_states_t0_to_H_B", "rewards_dreamed_t0_to_H_B", "continues_dreamed_t0_to_H_B",
                      "actions_dreamed_t0_to_H_B", "values_dreamed_t0_to_H_B",
                      "values_symlog_dreamed_logits_t0_to_HxB", "v_symlog_dreamed_ema_t0_to_H_B",
                      "dream_loss_weights_t0_to_H_B"):
                out[pre + "dream/" + k] = npf(dd[k])
            out[pre + "dream/z_indices_t1_to_H_B"] = dd["z_states_prior_t0_to_H_B"].numpy()[1:].argmax(-1).astype(np.int32)
            if cfg.discrete:
                out[pre + "dream/actions_ints_dreamed_t0_to_H_B"] = dd["actions_ints_dreamed_t0_to_H_B"].numpy().astype(np.int32)
            if cur:
                out[pre + "dream/rewards_intrinsic_t1_to_H_B"] = npf(dd["rewards_intrinsic_t1_to_H_B"])
                out[pre + "dream/z_predicted_probs_N_HxB"] = np.stack([npf(p) for p in dd["z_predicted_probs_N_HxB"]], 0)
            for k, v in r_ac.items():
                if k != "dream_data" and isinstance(v, torch.Tensor):
                    out[pre + "ac/" + k] = npf(v)
            # ---- gradients (the lists handed to tf.clip_by_global_norm: WM, actor, critic) ----------
            assert len(grads_seen) == len(models)   # the lists handed to tf.clip_by_global_norm: WM, actor, critic (, disagree)
            for (grp, model), gl in zip(models, grads_seen):
                for var, g in zip(model.trainable_variables, gl):
                    out[pre + "grad/" + by_id[id(var)]] = npf(g)
            out.update({pre + "param/" + n: npf(v) for n, v in vm.items()})
        # ---- dream_trajectory_with_burn_in (dreamer_model.py:306-413), called as run_experiment.py:629-644 does
        # (given actions) and in its default mode (actor's actions), on the twice-updated weights ----
        burn = cfg.T - H  # burn_in_T + horizon_H <= batch_length_T (run_experiment.py:141-144)
        start = {k: tf.repeat(v, cfg.B, axis=0) for k, v in dm.get_initial_state().items()}
        for mode, sampled in (("given", True), ("actor", False)):
            tfp.DRAW_LOG.clear()
            bd = dm.dream_trajectory_with_burn_in(
                start_states=dict(start), timesteps_burn_in=burn, timesteps_H=H,
                observations=sample["obs"][:, :burn], actions=sample["actions"][:, :burn + H] if sampled
                else sample["actions"][:, :burn], use_sampled_actions_in_dream=sampled, use_random_actions_in_dream=False)
            log = list(tfp.DRAW_LOG)
            pre = f"burn_{mode}/"
            # per burn-in step: initial-state draw (discarded) + posterior draw; per dream step: prior draw (+ actor draw)
            per = 1 if sampled else 2
            assert len(log) == 2 * burn + per * H, (len(log), burn, H)
            out[pre + "noise/u_burn"] = np.stack([log[2 * i + 1][1] for i in range(burn)], 0)            # [burn,B,Zc]
            out[pre + "noise/u_dyn"] = np.stack([log[2 * burn + per * j][1] for j in range(H)], 0)        # [H,B,Zc]
            if not sampled:
                out[pre + "noise/act"] = np.stack([log[2 * burn + 2 * j + 1][1] for j in range(H)], 0)    # [H,B(,A)]
            for k, v in bd.items():
                out[pre + k] = v.numpy().astype(np.int32) if "ints" in k else npf(v)
            out[pre + "z_indices_t1_to_H_B"] = bd["z_states_prior_t0_to_H_B"].numpy()[1:].argmax(-1).astype(np.int32)
        # ---- DreamerModel.forward_inference (dreamer_model.py:113-128), the acting step of env_runner_v2.py:274-282, on
        # the same weights: previous states taken from the last observe pass, a mixed is_first vector ----
        fwd = r_wm["WORLD_MODEL_forward_train_outs"]
        rows = np.arange(cfg.B) * cfg.T + 1          # (b, t=1) of the folded B x T rows
        prev = {"h": tf.convert_to_tensor(fwd["h_states_BxT"].numpy()[rows], dtype=dtype),
                "z": tf.convert_to_tensor(fwd["z_states_BxT"].numpy()[rows], dtype=dtype),
                "a": sample["actions"][:, 1]}
        first = tf.convert_to_tensor(np.array([1.0] + [0.0] * (cfg.B - 1)), dtype=dtype)
        tfp.DRAW_LOG.clear()
        act, st = dm.forward_inference(prev, sample["obs"][:, 2], first)
        log = list(tfp.DRAW_LOG)
        assert len(log) == 3 and log[0][1].shape == (1, Zc)   # initial-state draw (discarded), posterior z, action
        out["infer/noise/u_z"], out["infer/noise/act"] = log[1][1], log[2][1]
        out["infer/prev_h"], out["infer/prev_z"], out["infer/prev_a"] = npf(prev["h"]), npf(prev["z"]), npf(prev["a"])
        out["infer/obs"], out["infer/is_first"] = npf(sample["obs"][:, 2]), npf(first)
        out["infer/h"], out["infer/a"] = npf(st["h"]), npf(act)
        out["infer/z_indices"] = st["z"].numpy().argmax(-1).astype(np.int32)
    finally:
        tf.clip_by_global_norm = real_clip
    return out


def main():
    only = sys.argv[1:]   # optional case names; default: all
    for name, (kw, seed) in CASES.items():
        if only and name not in only:
            continue
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            out = run_case(name, kw, seed, dtype)
            path = os.path.join(HERE, f"ref_{name}_{tag}.npz")
            np.savez_compressed(path, **out)
            print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
