"""Invariants of the torch-CPU oracle: the reference pins none of this by tests,
so the checks are the in-graph asserts / documented properties of the reference."""
import math

import numpy as np
import pytest
import torch

from dreamer_v3_b200.spec import EngineConfig, init_params
from oracle.dreamer_oracle import Oracle, make_synthetic_sample


def run(cfg, mode="keras", dtype=torch.float32, seed=0):
    P = init_params(cfg, seed=seed, mode=mode)
    s = make_synthetic_sample(cfg, seed=seed + 1)
    o = Oracle(cfg, P, dtype=dtype)
    rng = np.random.default_rng(seed + 2)
    res = o.train_world_model_one_step(s, rng=rng, margin=1e-4)
    return o, s, rng, res


@pytest.mark.parametrize("kw", [dict(A=2, discrete=True), dict(A=3, discrete=False), dict(A=4, discrete=True, image_obs=True)])
def test_world_model_invariants(kw):
    cfg = EngineConfig.make("nano", B=3, T=6, H=4, **kw)
    o, s, rng, res = run(cfg)
    # train_one_step.py:71 tf.assert_equal(L_dyn, L_rep)
    assert float(res["WORLD_MODEL_L_dynamics"].detach()) == float(res["WORLD_MODEL_L_representation"].detach())
    assert float(res["WORLD_MODEL_L_dynamics_B_T"].min()) >= 1.0  # free bits
    # zero-initialised reward head => uniform buckets: E[r] = 0 and CE = log(255) (reward_predictor_layer.py:58-59)
    f = res["WORLD_MODEL_forward_train_outs"]
    assert float(f["rewards_BxT"].abs().max()) < 1e-5
    np.testing.assert_allclose(res["WORLD_MODEL_L_reward_B_T"].detach().numpy(), math.log(255), rtol=1e-5)
    # unimix: every class keeps >= 1% / Zk probability; rows sum to 1
    for k in ("z_posterior_probs_BxT", "z_prior_probs_BxT"):
        p = f[k]
        assert float(p.min()) >= 0.01 / cfg.Zk - 1e-7
        np.testing.assert_allclose(p.sum(-1).numpy(), 1.0, atol=1e-5)
    z = f["z_states_BxT"]
    # straight-through value (onehot + p) - p is one-hot up to one ulp (representation_layer.py:108-110)
    assert (((z - 0).abs() < 1e-6) | ((z - 1).abs() < 1e-6)).all() and ((z.sum(-1) - 1).abs() < 1e-5).all()
    assert all(torch.isfinite(g).all() for g in res["grads"].values())
    # initial state gradient only flows through tanh(initial_h); the heads of a zero-init model get zero grad
    assert float(res["grads"]["rew.mlp0.w"].abs().max()) == 0.0


def test_actor_critic_invariants_and_ema():
    cfg = EngineConfig.make("nano", B=3, T=6, H=4, A=2, discrete=True)
    o, s, rng, res = run(cfg, mode="random")
    f = res["WORLD_MODEL_forward_train_outs"]
    before = o.numpy_params()
    ac = o.train_actor_and_critic_one_step(f["h_states_BxT"], f["z_states_BxT"], s["is_terminated"], H=cfg.H, rng=rng)
    d = ac["dream_data"]
    w = d["dream_loss_weights_t0_to_H_B"]
    c = d["continues_dreamed_t0_to_H_B"]
    np.testing.assert_allclose(w[0].numpy(), c[0].numpy())                       # cumprod(gamma c)/gamma at t=0
    assert (w[1:] <= w[:-1] + 1e-7).all()                                        # weights never grow
    assert (c[0].numpy() == 1.0 - s["is_terminated"].reshape(-1)).all()          # dreamer_model.py:252-255
    assert ac["VALUE_TARGETS_H_B"].shape == (cfg.H, cfg.N)
    assert not math.isnan(o.ema_pct5) and o.ema_pct5 <= o.ema_pct95
    after = o.numpy_params()
    for n in o.group_names("critic"):
        e = "critic_ema." + n[len("critic."):]
        np.testing.assert_allclose(after[e], 0.98 * before[e] + 0.02 * after[n], rtol=1e-5, atol=1e-7)
    for n in o.group_names("world_model"):
        np.testing.assert_array_equal(before[n], after[n])  # AC step never touches the world model
    # second call: EMA percentiles move by (1 - decay)
    p5 = o.ema_pct5
    o.train_actor_and_critic_one_step(f["h_states_BxT"], f["z_states_BxT"], s["is_terminated"], H=cfg.H, rng=rng)
    assert o.ema_pct5 != p5


def test_box_actor_gradient_flows_through_dynamics():
    cfg = EngineConfig.make("nano", B=2, T=4, H=3, A=2, discrete=False)
    o, s, rng, res = run(cfg, mode="random")
    f = res["WORLD_MODEL_forward_train_outs"]
    ac = o.train_actor_and_critic_one_step(f["h_states_BxT"], f["z_states_BxT"], s["is_terminated"], H=cfg.H, rng=rng, apply=False)
    assert float(ac["grads"]["actor.out.w"].abs().max()) > 0
    assert float(ac["grads"]["actor.std_out.w"].abs().max()) > 0
    # -scaled value targets is the reinforce term for Box (actor_loss.py:57)
    np.testing.assert_allclose(ac["ACTOR_L_neglogp_reinforce_term_H_B"].detach().numpy(),
                               -ac["ACTOR_scaled_value_targets_H_B"].detach().numpy())


def test_fp32_vs_fp64_self_consistency():
    cfg = EngineConfig.make("micro", B=2, T=5, H=3, A=3, discrete=True)
    o32, s, rng, r32 = run(cfg, mode="random")
    o64 = Oracle(cfg, init_params(cfg, seed=0, mode="random"), dtype=torch.float64)
    r64 = o64.train_world_model_one_step(s, u_post=o32.last_noise["u_post"])
    np.testing.assert_array_equal(r64["WORLD_MODEL_forward_train_outs"]["z_indices_BxT"],
                                  r32["WORLD_MODEL_forward_train_outs"]["z_indices_BxT"])
    for n, g in r32["grads"].items():
        g64 = r64["grads"][n].float()
        assert float((g - g64).abs().max()) <= 2e-4 * float(g64.abs().max()) + 1e-9, n


def test_keras_adam_matches_numpy_restatement():
    from oracle import numerics as nx
    cfg = EngineConfig.make("nano", B=2, T=3, H=2, A=2, discrete=True)
    o, s, rng, res = run(cfg, mode="random")
    # redo step 1 of Adam by hand for one tensor
    P0 = init_params(cfg, seed=0, mode="random")
    n = "post.mlp0.w"
    gn = res["grad_global_norm"]
    g = res["grads"][n].numpy() * np.float32(1000.0 / max(gn, 1000.0))
    w, _, _ = nx.keras_adam_step(P0[n], g, np.zeros_like(g), np.zeros_like(g), 1, 1e-4, 1e-8)
    np.testing.assert_allclose(o.numpy_params()[n], w, rtol=0, atol=2e-7)


def test_oracle_reproduces_golden_fixture():
    """Regression pin: the committed fixture (tests/golden/make_golden.py) is reproduced by the oracle in fp32."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nano_vec_disc.npz"))
    cfg = EngineConfig.make("nano", A=2, discrete=True, B=3, T=5, H=4)
    P = {k[len("param/"):]: g[k] for k in g.files if k.startswith("param/")}
    sample = {k: g["sample/" + k] for k in ("obs", "actions", "rewards", "is_first", "is_terminated")}
    o = Oracle(cfg, P, dtype=torch.float32)
    res = o.train_world_model_one_step(sample, u_post=g["noise/u_post"], apply=False)
    np.testing.assert_array_equal(res["WORLD_MODEL_forward_train_outs"]["z_indices_BxT"], g["out/z_indices"])
    np.testing.assert_array_equal(res["reward_two_hot_k"], g["out/reward_two_hot_k"])
    np.testing.assert_allclose(res["WORLD_MODEL_L_total"].item(), g["out/L_total"], rtol=1e-5)
    for k in g.files:
        if k.startswith("grad/"):
            want = g[k]
            got = res["grads"][k[len("grad/"):]].numpy()
            assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max() + 1e-12, k


def test_gru_step_against_torch_grucell():
    """Independent check of the restated Keras GRU step (reset_after=True, gate order z, r, h): the same
    weights permuted into torch.nn.GRUCell's (r, z, n) layout must give the same h'."""
    import numpy as np
    import torch
    from dreamer_v3_b200.spec import EngineConfig, init_params
    from oracle.dreamer_oracle import Oracle
    cfg = EngineConfig.make("XXS", A=3, discrete=True, B=4, T=4, H=2)
    o = Oracle(cfg, init_params(cfg, seed=5, mode="random"), dtype=torch.float64)
    G, D = cfg.G, cfg.D
    rng = np.random.default_rng(0)
    z = torch.as_tensor(np.eye(cfg.Zk)[rng.integers(0, cfg.Zk, size=(5, cfg.Zc))])
    a = torch.as_tensor(np.eye(cfg.A)[rng.integers(0, cfg.A, size=5)])
    h = torch.as_tensor(rng.standard_normal((5, G)))
    got = o.sequence_model(z, a, h)
    u = o.mlp("seq.pre", torch.cat([z.reshape(5, -1), a], -1), 1)   # input of the GRU
    K, U, b = (o.P[k].detach() for k in ("seq.gru.kernel", "seq.gru.recurrent", "seq.gru.bias"))
    perm = lambda w: torch.cat([w[..., G:2 * G], w[..., :G], w[..., 2 * G:]], -1)   # (z, r, h) -> (r, z, n)  # noqa: E731
    cell = torch.nn.GRUCell(D, G, dtype=torch.float64)
    with torch.no_grad():
        cell.weight_ih.copy_(perm(K).T)
        cell.weight_hh.copy_(perm(U).T)
        cell.bias_ih.copy_(perm(b[0]))
        cell.bias_hh.copy_(perm(b[1]))
    want = cell(u.detach(), h)
    assert float((got - want).abs().max()) < 1e-12


def test_conv_layouts_against_tf_documented_semantics():
    """The oracle runs the Keras convolutions through torch ops on permuted kernels; this restates TensorFlow's documented
    definitions directly (explicit loops) to pin the layout / padding / no-flip conventions:
      Conv2D(k=4, s=2, 'same') on an even size: pad_total = (ceil(n/2) - 1) * 2 + 4 - n = 2, one pixel on each side;
          y[oy, ox, co] = sum x[2 oy - 1 + ky, 2 ox - 1 + kx, ci] * w[ky, kx, ci, co]
      Conv2DTranspose(k=4, s=2, 'same') = the gradient of that convolution w.r.t. its input, kernel [kh, kw, out, in]:
          y[2 iy - 1 + ky, 2 ix - 1 + kx, co] += x[iy, ix, ci] * w[ky, kx, co, ci]"""
    import numpy as np
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(1)
    n, ci, co = 6, 3, 2
    x = rng.standard_normal((1, n, n, ci))
    w = rng.standard_normal((4, 4, ci, co))
    want = np.zeros((1, n // 2, n // 2, co))
    for oy in range(n // 2):
        for ox in range(n // 2):
            for ky in range(4):
                for kx in range(4):
                    y_, x_ = 2 * oy - 1 + ky, 2 * ox - 1 + kx
                    if 0 <= y_ < n and 0 <= x_ < n:
                        want[0, oy, ox] += x[0, y_, x_] @ w[ky, kx]
    got = F.conv2d(torch.as_tensor(x).permute(0, 3, 1, 2), torch.as_tensor(w).permute(3, 2, 0, 1), stride=2, padding=1)
    assert np.abs(got.permute(0, 2, 3, 1).numpy() - want).max() < 1e-12      # the call Oracle.encoder makes
    wt = rng.standard_normal((4, 4, co, ci))    # Keras Conv2DTranspose kernel [kh, kw, out_ch, in_ch]
    xin = rng.standard_normal((1, 3, 3, ci))
    want_t = np.zeros((1, 6, 6, co))
    for iy in range(3):
        for ix in range(3):
            for ky in range(4):
                for kx in range(4):
                    y_, x_ = 2 * iy - 1 + ky, 2 * ix - 1 + kx
                    if 0 <= y_ < 6 and 0 <= x_ < 6:
                        want_t[0, y_, x_] += wt[ky, kx] @ xin[0, iy, ix]
    got_t = F.conv_transpose2d(torch.as_tensor(xin).permute(0, 3, 1, 2), torch.as_tensor(wt).permute(3, 2, 0, 1), stride=2, padding=1)
    assert np.abs(got_t.permute(0, 2, 3, 1).numpy() - want_t).max() < 1e-12  # the call Oracle.decoder makes


def test_forced_draws_reproduce_free_draws():
    """Oracle._draw forced mode (used by the bf16 GPU parity tests): forcing the indices the oracle itself drew from the same
    uniforms reproduces the run bit for bit with zero CDF violation and zero flips; forcing a shifted index is reported."""
    import numpy as np
    import torch
    from dreamer_v3_b200.spec import EngineConfig, init_params
    from oracle.dreamer_oracle import Oracle, make_synthetic_sample
    cfg = EngineConfig.make("nano", A=3, discrete=True, B=3, T=4, H=3)
    p = init_params(cfg, seed=1, mode="random")
    s = make_synthetic_sample(cfg, seed=2)
    o1 = Oracle(cfg, p, dtype=torch.float64)
    r1 = o1.train_world_model_one_step(s, rng=np.random.default_rng(0), apply=False)
    u_post = o1.last_noise["u_post"]
    f1 = r1["WORLD_MODEL_forward_train_outs"]
    a1 = o1.train_actor_and_critic_one_step(f1["h_states_BxT"], f1["z_states_BxT"], s["is_terminated"], H=cfg.H,
                                            rng=np.random.default_rng(1), apply=False)
    u_dyn, u_act = o1.last_noise["u_dyn"], o1.last_noise["u_act"]
    d1 = a1["dream_data"]
    o2 = Oracle(cfg, p, dtype=torch.float64)
    zi = f1["z_indices_BxT"].reshape(cfg.B, cfg.T, cfg.Zc)
    o2.set_forced({"u_post": [zi[:, t] for t in range(cfg.T)]})
    r2 = o2.train_world_model_one_step(s, u_post=u_post, apply=False)
    assert o2.forced_violation == {"u_post": 0.0} and o2.forced_flips == {"u_post": 0}
    assert float(r2["WORLD_MODEL_L_total"]) == float(r1["WORLD_MODEL_L_total"])
    o2.set_forced({"u_dyn": list(d1["z_indices_t1_to_H_B"]), "u_act": list(d1["actions_ints_dreamed_t0_to_H_B"].numpy()),
                   "continues": d1["continues_dreamed_t0_to_H_B"].numpy()[0:].reshape(-1) * 0 + (
                       o1.continue_predictor(torch.cat([d1["h_states_t0_to_H_B"].reshape(-1, cfg.G), d1["z_states_prior_t0_to_H_B"].reshape(-1, cfg.Z)], -1))[0].numpy())})
    a2 = o2.train_actor_and_critic_one_step(f1["h_states_BxT"], f1["z_states_BxT"], s["is_terminated"], H=cfg.H,
                                            u_dyn=u_dyn, act_noise=u_act, apply=False)
    assert all(v == 0.0 for v in o2.forced_violation.values()), o2.forced_violation
    assert all(v == 0 for v in o2.forced_flips.values()), o2.forced_flips
    assert float(a2["ACTOR_L_total"]) == float(a1["ACTOR_L_total"]) and float(a2["CRITIC_L_total"]) == float(a1["CRITIC_L_total"])
    # a wrong forced index is measured, not hidden
    o3 = Oracle(cfg, p, dtype=torch.float64)
    o3.set_forced({"u_post": [(zi[:, t] + 1) % cfg.Zk for t in range(cfg.T)]})
    o3.train_world_model_one_step(s, u_post=u_post, apply=False)
    assert o3.forced_flips["u_post"] >= cfg.B * cfg.Zc and o3.forced_violation["u_post"] > 0.0  # at least the whole first step
