"""The golden fixtures are the reference's own code run over `tests/golden/tfshim/` - so what they pin is only as good as the
shim's restatement of TensorFlow / Keras / tfp primitives. TensorFlow cannot be installed here; this file holds every shim
primitive with arithmetic of its own against an implementation written by someone else: `torch.nn` modules (GRUCell,
LayerNorm, Conv2d / ConvTranspose2d with the documented Keras layouts written out as explicit loops), `torch.distributions`
(the same definitions tfp uses), `torch.optim`-free closed-form Adam in numpy, `numpy.percentile`."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributions as td

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tfshim")


@pytest.fixture(scope="module")
def shim():
    """Imports the shim packages under their TensorFlow names for this module only."""
    names = ("tensorflow", "tensorflow.keras", "tensorflow_probability", "gymnasium", "tree")
    saved = {n: sys.modules.get(n) for n in list(sys.modules) if n.split(".")[0] in ("tensorflow", "tensorflow_probability", "gymnasium", "tree")}
    sys.path.insert(0, SHIM)
    try:
        tf = importlib.import_module("tensorflow")
        tfp = importlib.import_module("tensorflow_probability")
        tf.set_float(torch.float64)
        yield tf, tfp
    finally:
        sys.path.remove(SHIM)
        for n in list(sys.modules):
            if n.split(".")[0] in ("tensorflow", "tensorflow_probability", "gymnasium", "tree") and n not in saved:
                del sys.modules[n]
        del names


def _t(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def test_gru_against_torch_grucell(shim):
    """Keras GRU(reset_after=True), gate order z, r, h == torch.nn.GRUCell (gates r, z, n) with the blocks permuted."""
    tf, _ = shim
    rng = np.random.default_rng(0)
    D, G, B = 5, 7, 4
    layer = tf.keras.layers.GRU(G, time_major=True)
    x, h0 = _t(rng.standard_normal((1, B, D))), _t(rng.standard_normal((B, G)))
    layer(tf.convert_to_tensor(x), initial_state=tf.convert_to_tensor(h0))      # builds
    K, U, b = (_t(rng.standard_normal(tuple(v.shape))) for v in (layer.kernel, layer.recurrent_kernel, layer.bias))
    layer.kernel.assign(K); layer.recurrent_kernel.assign(U); layer.bias.assign(b)
    got = layer(tf.convert_to_tensor(x), initial_state=tf.convert_to_tensor(h0))
    cell = torch.nn.GRUCell(D, G).double()
    perm = lambda m: torch.cat([m[:, G:2 * G], m[:, :G], m[:, 2 * G:]], 1)       # (z, r, h) -> (r, z, n)  # noqa: E731
    with torch.no_grad():
        cell.weight_ih.copy_(perm(K).t()); cell.weight_hh.copy_(perm(U).t())
        cell.bias_ih.copy_(perm(b[0:1])[0]); cell.bias_hh.copy_(perm(b[1:2])[0])
    want = cell(x[0], h0)
    np.testing.assert_allclose(got.detach().numpy(), want.detach().numpy(), rtol=1e-12, atol=1e-12)


def test_layer_normalization_against_torch(shim):
    tf, _ = shim
    rng = np.random.default_rng(1)
    x = _t(rng.standard_normal((6, 3, 10)) * 2 + 1)
    ln = tf.keras.layers.LayerNormalization()
    ln(tf.convert_to_tensor(x))
    g, b = _t(rng.standard_normal(10)), _t(rng.standard_normal(10))
    ln.gamma.assign(g); ln.beta.assign(b)
    want = torch.nn.functional.layer_norm(x, (10,), g, b, eps=1e-3)              # biased variance, eps inside the root
    np.testing.assert_allclose(ln(tf.convert_to_tensor(x)).detach().numpy(), want.numpy(), rtol=1e-12, atol=1e-12)
    assert ln.epsilon == 1e-3


def _conv_same_k4s2_loops(x, k):
    """tf.nn.conv2d, NHWC, kernel [kh, kw, cin, cout], strides 2, padding 'SAME' on an even size: pad_total = k - s = 2 -> one
    pixel before, one after (TensorFlow's documented SAME rule), out[n, i, j, o] = sum x_pad[n, 2i + a, 2j + b, c] k[a, b, c, o]."""
    N, H, W, C = x.shape
    xp = np.zeros((N, H + 2, W + 2, C))
    xp[:, 1:-1, 1:-1] = x
    out = np.zeros((N, H // 2, W // 2, k.shape[3]))
    for i in range(H // 2):
        for j in range(W // 2):
            patch = xp[:, 2 * i:2 * i + 4, 2 * j:2 * j + 4, :]
            out[:, i, j, :] = np.einsum("nabc,abco->no", patch, k)
    return out


def test_conv2d_and_its_transpose_against_explicit_loops(shim):
    tf, _ = shim
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 8, 8, 3))
    conv = tf.keras.layers.Conv2D(5, kernel_size=4, strides=(2, 2), padding="same", use_bias=False)
    conv(tf.convert_to_tensor(_t(x)))
    k = rng.standard_normal((4, 4, 3, 5))
    conv.kernel.assign(_t(k))
    got = conv(tf.convert_to_tensor(_t(x))).detach().numpy()
    np.testing.assert_allclose(got, _conv_same_k4s2_loops(x, k), rtol=1e-12, atol=1e-12)
    # Conv2DTranspose(kernel [kh, kw, out, in]) is defined as the gradient of Conv2D w.r.t. its input: with the same array
    # as kernel it must be the adjoint, <conv(x), y> == <x, convT(y)>, and it doubles the spatial size
    convT = tf.keras.layers.Conv2DTranspose(3, kernel_size=4, strides=(2, 2), padding="same", use_bias=False)
    y = rng.standard_normal((2, 4, 4, 5))
    convT(tf.convert_to_tensor(_t(y)))
    assert tuple(convT.kernel.shape) == (4, 4, 3, 5)
    convT.kernel.assign(_t(k))
    back = convT(tf.convert_to_tensor(_t(y))).detach().numpy()
    assert back.shape == (2, 8, 8, 3)
    np.testing.assert_allclose((got * y).sum(), (x * back).sum(), rtol=1e-12)
    # and explicitly: every input pixel scatters its 4 x 4 stamp at stride 2, cropped by one pixel on each side
    full = np.zeros((2, 10, 10, 3))
    for i in range(4):
        for j in range(4):
            full[:, 2 * i:2 * i + 4, 2 * j:2 * j + 4, :] += np.einsum("no,abco->nabc", y[:, i, j, :], k)
    np.testing.assert_allclose(back, full[:, 1:-1, 1:-1], rtol=1e-12, atol=1e-12)


def test_dense_and_flatten(shim):
    tf, _ = shim
    rng = np.random.default_rng(3)
    x = _t(rng.standard_normal((4, 6)))
    d = tf.keras.layers.Dense(3)
    d(tf.convert_to_tensor(x))
    w, b = _t(rng.standard_normal((6, 3))), _t(rng.standard_normal(3))
    d.kernel.assign(w); d.bias.assign(b)
    np.testing.assert_allclose(d(tf.convert_to_tensor(x)).detach().numpy(), (x @ w + b).numpy(), rtol=1e-12)
    f = tf.keras.layers.Flatten()
    np.testing.assert_array_equal(f(tf.convert_to_tensor(_t(np.arange(24).reshape(2, 3, 4)))).numpy(), np.arange(24).reshape(2, 12))


def test_adam_is_keras2_adam(shim):
    """Keras-2 OptimizerV2 Adam: m, v EMAs, step lr sqrt(1 - b2^t) / (1 - b1^t) m / (sqrt(v) + eps) - eps is added to the
    un-corrected sqrt(v), which is NOT torch.optim.Adam's rule (that one divides sqrt(v) by sqrt(1 - b2^t) first)."""
    tf, _ = shim
    rng = np.random.default_rng(4)
    w0 = rng.standard_normal(7)
    var = tf.Variable(_t(w0))
    opt = tf.keras.optimizers.Adam(learning_rate=1e-2, epsilon=1e-5)
    w, m, v = w0.copy(), np.zeros(7), np.zeros(7)
    for t in range(1, 6):
        g = rng.standard_normal(7) * 10.0 ** rng.integers(-6, 1)
        opt.apply_gradients([(tf.convert_to_tensor(_t(g)), var)])
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        w = w - 1e-2 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-5)
        np.testing.assert_allclose(var.detach().numpy(), w, rtol=1e-12, atol=1e-15)
    ref = torch.nn.Parameter(_t(w0).clone())
    topt = torch.optim.Adam([ref], lr=1e-2, eps=1e-5)
    ref.grad = _t(np.full(7, 1e-6))
    topt.step()
    var2 = tf.Variable(_t(w0))
    tf.keras.optimizers.Adam(learning_rate=1e-2, epsilon=1e-5).apply_gradients([(tf.convert_to_tensor(_t(np.full(7, 1e-6))), var2)])
    assert np.abs(var2.detach().numpy() - ref.detach().numpy()).max() > 1e-4     # the two rules really differ for small gradients


def test_clip_by_global_norm(shim):
    tf, _ = shim
    rng = np.random.default_rng(5)
    gs = [_t(rng.standard_normal(s)) for s in ((3, 4), (5,), (2, 2, 2))]
    norm = float(np.sqrt(sum((g.numpy() ** 2).sum() for g in gs)))
    for clip in (0.5 * norm, 2.0 * norm):
        out, n = tf.clip_by_global_norm([tf.convert_to_tensor(g) for g in gs], clip)
        np.testing.assert_allclose(float(n), norm, rtol=1e-12)
        for o, g in zip(out, gs):
            np.testing.assert_allclose(o.detach().numpy(), g.numpy() * clip / max(norm, clip), rtol=1e-12)


def test_tfp_distributions_against_torch_distributions(shim):
    tf, tfp = shim
    tfd = tfp.distributions
    rng = np.random.default_rng(6)
    p = torch.softmax(_t(rng.standard_normal((5, 4, 6))), -1)
    q = torch.softmax(_t(rng.standard_normal((5, 4, 6))), -1)
    onehot = torch.nn.functional.one_hot(torch.as_tensor(rng.integers(0, 6, (5, 4))), 6).double()
    a = tfd.Independent(tfd.OneHotCategorical(probs=tf.convert_to_tensor(p)), reinterpreted_batch_ndims=1)
    b = tfd.Independent(tfd.OneHotCategorical(probs=tf.convert_to_tensor(q)), reinterpreted_batch_ndims=1)
    ta, tb = td.Independent(td.OneHotCategorical(probs=p), 1), td.Independent(td.OneHotCategorical(probs=q), 1)
    np.testing.assert_allclose(tfd.kl_divergence(a, b).numpy(), td.kl_divergence(ta, tb).numpy(), rtol=1e-12)
    np.testing.assert_allclose(a.log_prob(tf.convert_to_tensor(onehot)).numpy(), ta.log_prob(onehot).numpy(), rtol=1e-12)
    np.testing.assert_allclose(a.entropy().numpy(), ta.entropy().numpy(), rtol=1e-12)
    # un-normalised `logits=` (the actor passes log of unimixed probs; tfp re-normalises for entropy / log_prob)
    lg = _t(rng.standard_normal((3, 6)))
    c = tfd.OneHotCategorical(logits=tf.convert_to_tensor(lg))
    np.testing.assert_allclose(c.entropy().numpy(), td.OneHotCategorical(logits=lg).entropy().numpy(), rtol=1e-12)
    # Bernoulli(logits)
    lo = _t(rng.standard_normal(9) * 3)
    xb = _t(rng.integers(0, 2, 9))
    bern = tfd.Bernoulli(logits=tf.convert_to_tensor(lo))
    np.testing.assert_allclose(bern.log_prob(tf.convert_to_tensor(xb)).numpy(), td.Bernoulli(logits=lo).log_prob(xb).numpy(), rtol=1e-12)
    np.testing.assert_array_equal(bern.mode().numpy(), (lo > 0).numpy().astype(bern.mode().numpy().dtype))
    # Independent(Normal, 1)
    loc, scale, xs = _t(rng.standard_normal((4, 3))), _t(rng.uniform(0.1, 1.0, (4, 3))), _t(rng.standard_normal((4, 3)))
    n = tfd.Independent(tfd.Normal(tf.convert_to_tensor(loc), tf.convert_to_tensor(scale)), reinterpreted_batch_ndims=1)
    tn = td.Independent(td.Normal(loc, scale), 1)
    np.testing.assert_allclose(n.log_prob(tf.convert_to_tensor(xs)).numpy(), tn.log_prob(xs).numpy(), rtol=1e-12)
    np.testing.assert_allclose(n.entropy().numpy(), tn.entropy().numpy(), rtol=1e-12)
    # reparameterised sample = loc + scale * eps with the logged eps; categorical sample = inverse CDF of the logged uniform
    tfp.RNG = np.random.default_rng(7)
    tfp.DRAW_LOG.clear()
    s = tfd.Normal(tf.convert_to_tensor(loc), tf.convert_to_tensor(scale)).sample()
    np.testing.assert_allclose(s.numpy(), (loc + scale * _t(tfp.DRAW_LOG[-1][1])).numpy(), rtol=1e-12)
    z = tfd.OneHotCategorical(probs=tf.convert_to_tensor(p)).sample().numpy()
    u = tfp.DRAW_LOG[-1][1]
    cdf = np.cumsum(p.numpy(), -1)
    want = np.array([[int(np.searchsorted(cdf[i, j], u[i, j], side="right")) for j in range(4)] for i in range(5)])
    np.testing.assert_array_equal(z.argmax(-1), np.minimum(want, 5))
    # percentile, defaults (nearest)
    xv = _t(rng.standard_normal(1001))
    for qq in (5, 95):
        assert float(tfp.stats.percentile(tf.convert_to_tensor(xv), qq)) == float(np.percentile(xv.numpy(), qq, method="nearest"))


def test_tensors_are_immutable_under_augmented_assignment(shim):
    """tf.Tensor has no in-place arithmetic: `a += b` rebinds. critic_loss.py:117-120 relies on it (it adds the intrinsic
    rewards to a slice of dream_data["rewards_dreamed_t0_to_H_B"], which must stay the extrinsic prediction)."""
    tf, _ = shim
    r = tf.convert_to_tensor(_t(np.arange(6.0).reshape(3, 2)))
    tail = r[1:]
    tail += 100.0
    tail *= 2.0
    tail -= 1.0
    np.testing.assert_array_equal(r.numpy(), np.arange(6.0).reshape(3, 2))
    np.testing.assert_array_equal(tail.numpy(), (np.arange(6.0).reshape(3, 2)[1:] + 100.0) * 2.0 - 1.0)
    g = tf.stop_gradient(r)
    assert not g.requires_grad
