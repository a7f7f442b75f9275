"""Pins the oracle (and, on a GPU, the CUDA engine) against golden vectors produced by the REFERENCE'S OWN CODE
(tests/golden/make_reference_golden.py: /root/reference executed unmodified over the TF-API shim in
tests/golden/tfshim/).  The fixtures travel; /root/reference is not read here.

Two full updates per case: forward_train tensors, losses, dream tensors, lambda-return targets, actor / critic
losses, percentile EMA, every gradient, and the parameters after Adam / critic-EMA (t = 1 and t = 2).

Tolerances: integers (sampled category indices, action indices, continue flags) exact; float64 fixtures vs the
float64 oracle 1e-9 relative (max-norm; 1e-6 downstream of the float32 percentile EMA); float32 fixtures vs the float32 oracle 1e-4; CUDA engine (fp32 mode) vs the
float64 fixtures 1e-4 (BASELINE.json's fp32 tolerance).
"""
import ast
import os

import numpy as np
import pytest
import torch

from dreamer_v3_b200.spec import EngineConfig
from oracle.dreamer_oracle import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("vec_disc", "vec_box", "img_disc")
# use_curiosity=True (models/disagree_networks.py, losses/disagree_loss.py): 8 Plan2Explore disagree nets at nano
CUR_CASES = ("vec_disc_cur", "vec_box_cur")


def load(case, tag):
    z = np.load(os.path.join(GOLD, f"ref_{case}_{tag}.npz"))
    cfg = EngineConfig.make(**dict(ast.literal_eval(str(z["meta/cfg"]))))
    return z, cfg


def group(z, prefix):
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def tnp(x):
    return x.detach().numpy() if torch.is_tensor(x) else np.asarray(x)


WM_KEYS = ("WORLD_MODEL_L_decoder_B_T", "WORLD_MODEL_L_reward_B_T", "WORLD_MODEL_L_continue_B_T",
           "WORLD_MODEL_L_prediction_B_T", "WORLD_MODEL_L_dynamics_B_T", "WORLD_MODEL_L_representation_B_T",
           "WORLD_MODEL_L_total_B_T", "WORLD_MODEL_L_total", "WORLD_MODEL_L_decoder", "WORLD_MODEL_L_reward",
           "WORLD_MODEL_L_continue", "WORLD_MODEL_L_prediction", "WORLD_MODEL_L_dynamics",
           "WORLD_MODEL_L_representation", "WORLD_MODEL_gradients_maxabs",
           "WORLD_MODEL_gradients_clipped_by_glob_norm_maxabs")
AC_KEYS = ("VALUE_TARGETS_H_B", "VALUE_TARGETS_symlog_H_B", "CRITIC_L_total", "CRITIC_L_total_H_B",
           "CRITIC_L_neg_logp_of_value_targets_H_B", "CRITIC_L_neg_logp_of_value_targets",
           "CRITIC_L_slow_critic_regularization_H_B", "CRITIC_L_slow_critic_regularization", "ACTOR_L_total_H_B",
           "ACTOR_L_total", "ACTOR_logp_actions_dreamed_H_B", "ACTOR_scaled_value_targets_H_B",
           "ACTOR_value_targets_pct95_ema", "ACTOR_value_targets_pct5_ema", "ACTOR_action_entropy_H_B",
           "ACTOR_action_entropy", "ACTOR_L_neglogp_reinforce_term_H_B", "ACTOR_L_neglogp_reinforce_term",
           "ACTOR_L_neg_entropy_term_H_B", "ACTOR_L_neg_entropy_term", "ACTOR_gradients_maxabs",
           "CRITIC_gradients_maxabs")
DREAM_KEYS = ("h_states_t0_to_H_B", "rewards_dreamed_t0_to_H_B", "continues_dreamed_t0_to_H_B",
              "actions_dreamed_t0_to_H_B", "values_dreamed_t0_to_H_B", "values_symlog_dreamed_logits_t0_to_HxB",
              "v_symlog_dreamed_ema_t0_to_H_B", "dream_loss_weights_t0_to_H_B")


CUR_AC_KEYS = ("DISAGREE_L_total", "DISAGREE_intrinsic_rewards_H_B", "DISAGREE_intrinsic_rewards", "DISAGREE_gradients_maxabs",
               "DISAGREE_gradients_clipped_by_glob_norm_maxabs")


@pytest.mark.parametrize("tag,tol", (("f64", 1e-9), ("f32", 1e-4)))
@pytest.mark.parametrize("case", CASES + CUR_CASES)
def test_oracle_matches_reference_code(case, tag, tol):
    z, cfg = load(case, tag)
    dtype = torch.float64 if tag == "f64" else torch.float32
    o = Oracle(cfg, group(z, "param0/"), dtype=dtype)
    if tag == "f64":
        o.th_float = np.float64  # the float64 fixture ran symlog / two_hot in double as well
    sample = group(z, "sample/")
    err, exact = {}, {}
    for step in (1, 2):
        g = group(z, f"s{step}/")
        res = o.train_world_model_one_step(sample, u_post=g["noise/u_post"])
        fwd = res["WORLD_MODEL_forward_train_outs"]
        exact[f"s{step}.z_indices"] = int((fwd["z_indices_BxT"] != g["wm/z_indices_BxT"]).sum())
        for k_o, k_g in (("obs_distribution_BxT.loc", "obs_loc"), ("continue_distribution_BxT.logits", "continue_logits"),
                         ("reward_logits_BxT", "reward_logits_BxT"), ("rewards_BxT", "rewards_BxT"),
                         ("continues_BxT", "continues_BxT"), ("z_posterior_probs_BxT", "z_posterior_probs_BxT"),
                         ("z_prior_probs_BxT", "z_prior_probs_BxT"), ("h_states_BxT", "h_states_BxT")):
            err[f"s{step}.wm.{k_g}"] = rel(tnp(fwd[k_o]), g["wm/" + k_g])
        for k in WM_KEYS:
            err[f"s{step}.{k}"] = rel(tnp(res[k]), g["wm/" + k])
        for n, gr in res["grads"].items():
            err[f"s{step}.grad.{n}"] = rel(tnp(gr), g["grad/" + n])
        ac = o.train_actor_and_critic_one_step(fwd["h_states_BxT"], fwd["z_states_BxT"], sample["is_terminated"],
                                               H=cfg.H, u_dyn=g["noise/u_dyn"], act_noise=g["noise/act"])
        dd = ac["dream_data"]
        exact[f"s{step}.dream_z_indices"] = int((dd["z_indices_t1_to_H_B"] != g["dream/z_indices_t1_to_H_B"]).sum())
        if cfg.discrete:
            exact[f"s{step}.action_ints"] = int(
                (tnp(dd["actions_ints_dreamed_t0_to_H_B"]) != g["dream/actions_ints_dreamed_t0_to_H_B"]).sum())
        exact[f"s{step}.continue_flags"] = int(
            (tnp(dd["continues_dreamed_t0_to_H_B"]) != g["dream/continues_dreamed_t0_to_H_B"]).sum())
        for k in DREAM_KEYS:
            err[f"s{step}.dream.{k}"] = rel(tnp(dd[k]), g["dream/" + k])
        if cfg.num_curiosity_nets > 0:
            err[f"s{step}.dream.rewards_intrinsic"] = rel(tnp(dd["rewards_intrinsic_t1_to_H_B"]), g["dream/rewards_intrinsic_t1_to_H_B"])
            err[f"s{step}.dream.z_predicted_probs"] = rel(np.stack([tnp(p) for p in dd["z_predicted_probs_N_HxB"]], 0),
                                                          g["dream/z_predicted_probs_N_HxB"])
            assert any(n.startswith("disagree.") for n in ac["grads"])
        for k in AC_KEYS + (CUR_AC_KEYS if cfg.num_curiosity_nets > 0 else ()):
            err[f"s{step}.{k}"] = rel(tnp(ac[k]), g["ac/" + k])
        for n, gr in ac["grads"].items():
            err[f"s{step}.grad.{n}"] = rel(tnp(gr), g["grad/" + n])
        after = o.numpy_params() if tag == "f32" else {k: v.detach().numpy() for k, v in o.P.items()}
        for n, v in after.items():
            err[f"s{step}.param.{n}"] = rel(v, g["param/" + n])
    assert all(v == 0 for v in exact.values()), exact
    # the oracle keeps the percentile EMA in float32 like the reference's tf.Variable (actor_network.py:31-36), so
    # in the float64 comparison everything downstream of it (actor loss terms, actor gradients) carries ~1e-7
    loose = lambda k: tag == "f64" and ("ACTOR" in k or "actor" in k)  # noqa: E731
    bad = {k: v for k, v in err.items() if not v <= (1e-6 if loose(k) else tol)}
    assert not bad, dict(sorted(bad.items(), key=lambda kv: -kv[1])[:12])
    assert len(err) > 150


def test_reference_fixture_pins_the_wraparound_action():
    """world_model.py:253 reads actions[t-1] at t=0, i.e. the LAST timestep's action, for rows whose first step is
    not `is_first`.  The fixture has such a row; zeroing that action must change h at t=0 (so the pin is live)."""
    z, cfg = load("vec_disc", "f64")
    sample = group(z, "sample/")
    assert sample["is_first"][0, 0] == 0.0
    g = group(z, "s1/")
    o = Oracle(cfg, group(z, "param0/"), dtype=torch.float64)
    h = tnp(o.forward_train(sample["obs"], sample["actions"], sample["is_first"], u_post=g["noise/u_post"])["h_states_BxT"])
    assert rel(h, g["wm/h_states_BxT"]) < 1e-9
    s2 = dict(sample)
    s2["actions"] = sample["actions"].copy()
    s2["actions"][0, -1] = 0.0
    h2 = tnp(o.forward_train(s2["obs"], s2["actions"], s2["is_first"], u_post=g["noise/u_post"])["h_states_BxT"])
    assert np.abs(h2[0] - h[0]).max() > 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + CUR_CASES)
def test_cuda_engine_matches_reference_code(case):
    """The CUDA engine (fp32 mode) against the vectors the reference's own code produced in float32, fed the very
    noise the reference consumed: drawn category / action indices, continue flags and two-hot bucket indices
    bit-exact; every tensor, loss and gradient of two consecutive updates within 1e-4 relative (max-norm); Adam /
    EMA results from identical gradients within 1e-3 of the update size (fp32 weight resolution removed)."""
    from _parity import ParityRun
    z, cfg = load(case, "f32")
    run = ParityRun(cfg, steps=2, fixture=z)
    try:
        assert all(v == 0 for v in run.exact.values()), run.exact
        bad = {k: v for k, v in run.report.items() if not v <= 1e-4 and not k.startswith("adam_delta.")}
        assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:20]
        bad = {k: v for k, v in run.report.items() if k.startswith("adam_delta.") and not v <= 1e-3}
        assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:20]
        assert len(run.report) > 100
    finally:
        run.close()


def _burn_inputs(z, cfg, mode):
    g = group(z, f"burn_{mode}/")
    sample = group(z, "sample/")
    burn = cfg.T - cfg.H
    acts = sample["actions"][:, :burn + cfg.H] if mode == "given" else sample["actions"][:, :burn]
    return g, sample, burn, acts


BURN_FLOAT_KEYS = ("h_states_t0_to_H_B", "rewards_dreamed_t0_to_H_B")


@pytest.mark.parametrize("mode", ("given", "actor"))
@pytest.mark.parametrize("case", CASES)
def test_oracle_burn_in_dream_matches_reference_code(case, mode):
    """dream_trajectory_with_burn_in (dreamer_model.py:306-413) on the twice-updated weights of the fixture."""
    z, cfg = load(case, "f64")
    g, sample, burn, acts = _burn_inputs(z, cfg, mode)
    o = Oracle(cfg, group(z, "s2/param/"), dtype=torch.float64)
    init = o.get_initial_state()
    start = {"h": init["h"].detach().repeat(cfg.B, 1), "z": init["z"].repeat(cfg.B, 1, 1),
             "a": torch.zeros(cfg.B, cfg.A, dtype=torch.float64)}
    out = o.dream_trajectory_with_burn_in(start, burn, cfg.H, sample["obs"][:, :burn], acts, g["noise/u_burn"],
                                          g["noise/u_dyn"], g.get("noise/act"), use_sampled_actions_in_dream=(mode == "given"))
    akey = "actions_sampled_t0_to_H_B" if mode == "given" else "actions_dreamed_t0_to_H_B"
    assert (out["z_indices_t1_to_H_B"] != g["z_indices_t1_to_H_B"]).sum() == 0
    assert (tnp(out["continues_dreamed_t0_to_H_B"]) != g["continues_dreamed_t0_to_H_B"]).sum() == 0
    if cfg.discrete:
        ik = akey.replace("actions_", "actions_ints_", 1)
        assert (tnp(out[ik]) != g[ik]).sum() == 0
    for k in BURN_FLOAT_KEYS + (akey,):
        assert rel(tnp(out[k]), g[k]) <= 1e-9, k


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ("given", "actor"))
@pytest.mark.parametrize("case", CASES)
def test_cuda_burn_in_dream_matches_reference_code(case, mode):
    """DreamerModel.dream_trajectory_with_burn_in of the drop-in API (forward_inference steps + the given-action
    prior rollout kernels) against the float32 vectors of the reference's own code, fed its noise."""
    from dreamer_v3_b200.models.components import CNNAtari, ConvTransposeAtari, MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    z, cfg = load(case, "f32")
    g, sample, burn, acts = _burn_inputs(z, cfg, mode)
    space = Discrete(cfg.A) if cfg.discrete else Box(-1.0, 1.0, (cfg.A,))
    if cfg.image_obs:
        enc, dec = CNNAtari(model_dimension="nano"), ConvTransposeAtari(model_dimension="nano", gray_scaled=False)
    else:
        enc = MLP(model_dimension="nano")
        dec = VectorDecoder(model_dimension="nano", observation_space=Box(-10, 10, (cfg.obs_dim,)))
    wm = WorldModel(model_dimension="nano", action_space=space, batch_length_T=cfg.T, encoder=enc, decoder=dec,
                    symlog_obs=cfg.symlog_obs)
    dm = DreamerModel(model_dimension="nano", action_space=space, batch_size_B=cfg.B, batch_length_T=cfg.T,
                      horizon_H=cfg.H, world_model=wm)
    e = dm.engine
    try:
        e.load_params(group(z, "s2/param/"))
        start = {k: v.repeat_interleave(cfg.B, 0) for k, v in dm.get_initial_state().items()}
        # noise: posterior draws of the burn-in steps, then prior (and actor) draws of the rollout
        u_dyn = np.full((cfg.H, cfg.N, cfg.Zc), 0.5, np.float32)
        u_dyn[:, :cfg.B] = g["noise/u_dyn"]
        act = np.full((cfg.H + 1, cfg.N) + (() if cfg.discrete else (cfg.A,)), 0.5, np.float32)
        if mode == "actor":
            act[1:, :cfg.B] = g["noise/act"]
        wm.set_noise(u_dyn=u_dyn, act_noise=act)
        steps = iter(range(burn))
        orig = wm.forward_inference

        def with_noise(states, obs, first):
            wm.set_inference_noise(g["noise/u_burn"][next(steps)], np.full((cfg.B,) + (() if cfg.discrete else (cfg.A,)), 0.5, np.float32))
            return orig(states, obs, first)
        wm.forward_inference = with_noise
        out = dm.dream_trajectory_with_burn_in(
            start_states=start, timesteps_burn_in=burn, timesteps_H=cfg.H, observations=sample["obs"][:, :burn],
            actions=acts, use_sampled_actions_in_dream=(mode == "given"))
        akey = "actions_sampled_t0_to_H_B" if mode == "given" else "actions_dreamed_t0_to_H_B"
        zi = out["z_states_prior_t0_to_H_B"][1:].argmax(-1).cpu().numpy()
        assert (zi != g["z_indices_t1_to_H_B"]).sum() == 0
        assert (out["continues_dreamed_t0_to_H_B"].cpu().numpy() != g["continues_dreamed_t0_to_H_B"]).sum() == 0
        if cfg.discrete:
            ik = akey.replace("actions_", "actions_ints_", 1)
            assert (out[ik].cpu().numpy() != g[ik]).sum() == 0
        for k in BURN_FLOAT_KEYS + (akey,):
            assert rel(out[k].cpu().numpy(), g[k]) <= 1e-4, k
    finally:
        e.close()


@pytest.mark.parametrize("case", CASES)
def test_oracle_acting_step_matches_reference_code(case):
    """DreamerModel.forward_inference (dreamer_model.py:113-128 -> world_model.py:141-180 + actor) on the fixture's
    twice-updated weights, with a mixed is_first vector."""
    z, cfg = load(case, "f64")
    g = group(z, "infer/")
    o = Oracle(cfg, group(z, "s2/param/"), dtype=torch.float64)
    out = o.forward_inference(g["prev_h"], g["prev_z"], g["prev_a"], g["obs"], g["is_first"], u_z=g["noise/u_z"],
                              act_noise=g["noise/act"])
    assert (out["z_idx"] != g["z_indices"]).sum() == 0
    assert rel(tnp(out["h"]), g["h"]) <= 1e-9
    assert rel(tnp(out["a"]), g["a"]) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_acting_step_matches_reference_code(case):
    from dreamer_v3_b200.engine import Engine
    z, cfg = load(case, "f32")
    g = group(z, "infer/")
    e = Engine(cfg)
    try:
        e.load_params(group(z, "s2/param/"))
        rows = cfg.B
        for name, key in (("inf.prev_h", "prev_h"), ("inf.prev_z", "prev_z"), ("inf.prev_a", "prev_a"), ("inf.obs", "obs"),
                          ("inf.is_first", "is_first"), ("inf.u_z", "noise/u_z"), ("inf.act_noise", "noise/act")):
            dst = e.tensor(name)
            dst.copy_(torch.as_tensor(g[key], dtype=torch.float32).reshape(dst.shape))
        e.forward_inference(rows, 0)
        hz = e.tensor("inf.hz").cpu().numpy()
        assert (e.tensor("inf.z_indices").cpu().numpy() != g["z_indices"]).sum() == 0
        assert rel(hz[:, :cfg.G], g["h"]) <= 1e-4
        assert rel(e.tensor("inf.a").cpu().numpy(), g["a"]) <= 1e-4
        assert (hz[:, cfg.G:].reshape(rows, cfg.Zc, cfg.Zk).argmax(-1) != g["z_indices"]).sum() == 0
    finally:
        e.close()
