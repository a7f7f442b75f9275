"""Independent checks of the third-party arithmetic the oracle restates (Keras LayerNormalization,
tfp distributions, tfp.stats.percentile). TensorFlow / tfp cannot be installed here, so each restated formula is held
against an implementation written by someone else - `torch.distributions` (same definitions as tfp's: KL of
Independent(OneHotCategorical), Bernoulli / Normal log-prob and entropy), `numpy.percentile(method="nearest")`, and the
LayerNormalization formula of the Keras documentation stated in plain numpy - on tensors the oracle itself produced.
(GRU cell vs torch.nn.GRUCell, convolutions and Keras Adam have their checks in tests/test_oracle_model.py.)"""
import math

import numpy as np
import pytest
import torch
import torch.distributions as td

import oracle.numerics as nx
from dreamer_v3_b200.spec import EngineConfig, init_params
from oracle.dreamer_oracle import LN_EPS, Oracle, make_synthetic_sample


def _run(discrete, dtype=torch.float64, seed=0, **kw):
    cfg = EngineConfig.make("nano", B=3, T=6, H=4, A=3 if discrete else 2, discrete=discrete, **kw)
    o = Oracle(cfg, init_params(cfg, seed=seed, mode="random"), dtype=dtype)
    s = make_synthetic_sample(cfg, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    res = o.train_world_model_one_step(s, rng=rng, margin=1e-4, apply=False)
    return cfg, o, s, rng, res


def test_layer_normalization_is_keras_formula():
    """keras.layers.LayerNormalization() defaults: axis -1, epsilon 1e-3 INSIDE the square root, biased variance,
    y = (x - mean) / sqrt(var + eps) * gamma + beta (components/mlp.py:58-61 relies on them)."""
    cfg, o, *_ = _run(True)
    x = torch.randn(7, cfg.G + cfg.Z, dtype=torch.float64)
    w, g, b = (o.P[f"rew.mlp0.{k}"].detach().numpy() for k in ("w", "g", "b"))
    pre = x.numpy() @ w
    mean = pre.mean(-1, keepdims=True)
    var = ((pre - mean) ** 2).mean(-1, keepdims=True)           # biased (divide by D)
    ln = (pre - mean) / np.sqrt(var + 1e-3) * g + b
    want = ln / (1.0 + np.exp(-ln))                              # SiLU
    got = o.mlp("rew", x, 1).detach().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    assert LN_EPS == 1e-3
    # the unbiased variance or an epsilon outside the root would be visible at this tolerance
    var_u = var * pre.shape[-1] / (pre.shape[-1] - 1)
    assert np.abs((pre - mean) / np.sqrt(var_u + 1e-3) * g + b - ln).max() > 1e-4


def test_kl_terms_against_torch_distributions():
    """world_model_losses.py:88-151: KL(Independent(OneHotCategorical(probs=post), 1) || ...prior) summed over the Zc
    categoricals, free bits max(1, .), both stop-gradient variants equal in value."""
    cfg, o, s, rng, res = _run(True)
    f = res["WORLD_MODEL_forward_train_outs"]
    post, prior = f["z_posterior_probs_BxT"].detach(), f["z_prior_probs_BxT"].detach()
    kl = td.kl_divergence(td.Independent(td.OneHotCategorical(probs=post), 1), td.Independent(td.OneHotCategorical(probs=prior), 1))
    want = torch.clamp(kl, min=1.0).reshape(cfg.B, cfg.T)
    np.testing.assert_allclose(res["WORLD_MODEL_L_dynamics_B_T"].detach().numpy(), want.numpy(), rtol=1e-12)
    np.testing.assert_allclose(res["WORLD_MODEL_L_representation_B_T"].detach().numpy(), want.numpy(), rtol=1e-12)
    assert float((kl > 1.0).float().mean()) > 0  # the un-clipped branch is exercised


def test_continue_loss_against_torch_bernoulli():
    """continue_predictor.py:41-46 + world_model_losses.py:73-74: -Bernoulli(logits).log_prob(1 - is_terminated); mode = logit > 0."""
    cfg, o, s, rng, res = _run(True)
    f = res["WORLD_MODEL_forward_train_outs"]
    logits = f["continue_distribution_BxT.logits"].detach().reshape(-1)
    x = torch.as_tensor(1.0 - s["is_terminated"].reshape(-1), dtype=torch.float64)
    d = td.Bernoulli(logits=logits)
    np.testing.assert_allclose(res["WORLD_MODEL_L_continue_B_T"].detach().numpy().reshape(-1), (-d.log_prob(x)).numpy(), rtol=1e-12)
    np.testing.assert_array_equal(f["continues_BxT"].detach().numpy().reshape(-1), (logits > 0).numpy().astype(np.float64))


@pytest.mark.parametrize("discrete", (True, False))
def test_actor_terms_against_torch_distributions(discrete):
    """actor_loss.py:37-74: log-prob and entropy of the dreamed actions under OneHotCategorical(logits=log unimixed probs)
    (Discrete) / Independent(Normal(tanh mu, std), 1) (Box, actor_network.py:102-114)."""
    cfg, o, s, rng, res = _run(discrete)
    f = res["WORLD_MODEL_forward_train_outs"]
    ac = o.train_actor_and_critic_one_step(f["h_states_BxT"], f["z_states_BxT"], s["is_terminated"], H=cfg.H, rng=rng, apply=False)
    d = ac["dream_data"]
    dists = d["actions_dreamed_distributions_t0_to_H_B"][:-1]
    actions = d["actions_dreamed_t0_to_H_B"].detach()[:-1]
    logp, ent = [], []
    for i, di in enumerate(dists):
        if discrete:
            dist = td.OneHotCategorical(logits=torch.log(di["probs"].detach()))
            logp.append(dist.log_prob(torch.round(actions[i])))
        else:
            dist = td.Independent(td.Normal(di["loc"].detach(), di["std"].detach()), 1)
            logp.append(dist.log_prob(actions[i]))
        ent.append(dist.entropy())
    np.testing.assert_allclose(ac["ACTOR_logp_actions_dreamed_H_B"].detach().numpy(), torch.stack(logp).numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ac["ACTOR_action_entropy_H_B"].detach().numpy(), torch.stack(ent).numpy(), rtol=1e-10)
    if not discrete:  # std = 0.9 sigmoid(s + 2) + 0.1 stays inside [0.1, 1.0]
        std = torch.stack([di["std"] for di in dists])
        assert float(std.detach().min()) >= 0.1 and float(std.detach().max()) <= 1.0


def test_reward_loss_is_categorical_cross_entropy_of_the_two_hot_target():
    """world_model_losses.py:45-59: -sum(log_softmax(logits) * two_hot(symlog(r))) == the cross entropy torch computes from
    the same soft target; expected reward = softmax . linspace(-20, 20, 255) (reward_predictor_layer.py:84-99)."""
    cfg, o, s, rng, res = _run(True)
    f = res["WORLD_MODEL_forward_train_outs"]
    logits = f["reward_logits_BxT"].detach()
    th = torch.as_tensor(nx.two_hot(nx.symlog(s["rewards"].reshape(-1).astype(np.float32))), dtype=torch.float64)
    np.testing.assert_allclose(th.sum(-1).numpy(), 1.0, atol=1e-6)
    want = torch.nn.functional.cross_entropy(logits, th, reduction="none")
    np.testing.assert_allclose(res["WORLD_MODEL_L_reward_B_T"].detach().numpy().reshape(-1), want.numpy(), rtol=1e-6)
    e = (td.Categorical(logits=logits).probs * torch.linspace(-20.0, 20.0, 255, dtype=torch.float64)).sum(-1)
    np.testing.assert_allclose(f["rewards_BxT"].detach().numpy().reshape(-1), e.numpy(), rtol=1e-9, atol=1e-12)


def test_percentile_nearest_against_numpy():
    """tfp.stats.percentile(x, q) defaults (interpolation='nearest') vs numpy's independent implementation, over sizes
    where (n - 1) q / 100 is not half-way (there the two libraries' tie rules may differ and tfp's is the contract)."""
    rng = np.random.default_rng(0)
    checked = 0
    for n in list(range(2, 60)) + [255, 1000, 15360, 16384]:
        x = rng.standard_normal(n).astype(np.float32)
        for q in (5, 50, 95):
            pos = (n - 1) * q / 100.0
            if abs(pos - math.floor(pos) - 0.5) < 1e-9:
                continue
            assert nx.percentile_nearest(x, q) == np.percentile(x, q, method="nearest"), (n, q)
            checked += 1
    assert checked > 140
    # the benchmark's own case: n = 15 * 1024 dreamed targets -> ascending indices 768 and 14591
    s = np.sort(rng.standard_normal(15360).astype(np.float32))
    assert nx.percentile_desc_index(15360, 5) == 15359 - 768 and nx.percentile_desc_index(15360, 95) == 15359 - 14591


def test_sampling_frequencies_follow_the_unimixed_probabilities():
    """OneHotCategorical(logits=log p).sample() restated as an inverse-CDF draw on explicit uniforms: empirical class
    frequencies over many uniforms must match p (chi-square, 31 degrees of freedom)."""
    rng = np.random.default_rng(3)
    logits = rng.standard_normal(32)
    p = np.exp(logits) / np.exp(logits).sum()
    p = (0.99 * p + 0.01 / 32).astype(np.float32)        # representation_layer.py:87
    n = 200000
    idx = nx.sample_categorical(np.broadcast_to(p, (n, 32)), rng.random(n, dtype=np.float32))
    counts = np.bincount(idx, minlength=32)
    chi2 = float(((counts - n * p) ** 2 / (n * p)).sum())
    assert chi2 < 70.0, chi2      # P(chi2_31 > 70) < 1e-4
    assert counts.min() > 0
