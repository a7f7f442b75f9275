"""Pins oracle/numerics.py against the reference's own known-answer vectors
(utils/two_hot.py:23-37 docstring; utils/two_hot.py:81-93 __main__ inputs with restated
outputs from SURVEY.md section 8c; losses/critic_loss.py:142-168 example inputs)."""
import numpy as np

from oracle import numerics as nx


def test_two_hot_docstring_vectors():
    # utils/two_hot.py:23-31
    np.testing.assert_allclose(
        nx.two_hot([2.5], 11, -5.0, 5.0)[0],
        [0, 0, 0, 0, 0, 0, 0, 0.5, 0.5, 0, 0], atol=1e-6)
    # utils/two_hot.py:32-37
    np.testing.assert_allclose(nx.two_hot([0.1], 5, -1.0, 1.0)[0], [0, 0, 0.8, 0.2, 0], atol=1e-6)


def test_two_hot_main_inputs():
    # utils/two_hot.py:84 boundary value: k == kp1 -> kp1 moved up, all weight on k
    out = nx.two_hot([0.0], 10, -5.0, 5.0)[0]
    np.testing.assert_allclose(out[[4, 5]], [0.5, 0.5], atol=1e-6)
    assert np.isfinite(out).all() and abs(out.sum() - 1) < 1e-6
    # utils/two_hot.py:86 out-of-range values are clipped
    k, kp1, wk, wkp1 = nx.two_hot_indices([20.1, 50.0, 150.0, -20.00001])
    assert k.tolist() == [254, 254, 254, 0]
    assert kp1.tolist() == [253, 253, 253, 1]
    np.testing.assert_allclose(wk, [1, 1, 1, 1], atol=1e-6)
    # utils/two_hot.py:89-93
    o = nx.two_hot([2.5, 0.1], 10, -5.0, 5.0)
    np.testing.assert_allclose(o[0, [6, 7]], [0.25, 0.75], atol=1e-5)
    np.testing.assert_allclose(o[1, [4, 5]], [0.41, 0.59], atol=1e-5)
    np.testing.assert_allclose(nx.two_hot([0.1], 4, -1.0, 1.0)[0, [1, 2]], [0.35, 0.65], atol=1e-5)
    o = nx.two_hot([-0.5, -1.2], 9, -6.0, 3.0)
    np.testing.assert_allclose(o[0, [4, 5]], [1 / 9, 8 / 9], atol=1e-5)
    np.testing.assert_allclose(o[1, [4, 5]], [0.73333, 0.26667], atol=1e-4)


def test_two_hot_255_buckets():
    k, kp1, wk, _ = nx.two_hot_indices([0.0, 1.0, -1.0])
    assert k.tolist() == [127, 133, 120]
    assert kp1.tolist() == [128, 134, 121]
    np.testing.assert_allclose(wk, [1.0, 0.649997, 0.350003], atol=2e-5)
    th = nx.two_hot(np.linspace(-25, 25, 1001))
    np.testing.assert_allclose(th.sum(-1), 1.0, atol=1e-5)
    assert ((th != 0).sum(-1) <= 2).all()
    # decode: sum(w * bucket) == clipped value
    buckets = np.linspace(-20, 20, 255)
    np.testing.assert_allclose(th @ buckets, np.clip(np.linspace(-25, 25, 1001), -20, 20), atol=2e-4)


def test_symlog_roundtrip():
    x = np.array([-1e6, -3.5, -1.0, -1e-3, 0.0, 1e-3, 1.0, 3.5, 1e6], np.float32)
    y = nx.symlog(x)
    np.testing.assert_allclose(nx.inverse_symlog(y), x, rtol=2e-5, atol=1e-7)
    assert y.dtype == np.float32 and y[4] == 0.0


def test_value_targets_example():
    # inputs of losses/critic_loss.py:142-160 (the reference call omits a required kwarg
    # and only prints; answers restated in SURVEY.md section 4)
    r = np.array([[99.0], [1.0], [2.0], [3.0], [4.0], [5.0]])
    c = np.array([[1.0], [1.0], [0.0], [1.0], [1.0], [1.0]])
    v = np.array([[3.0], [2.0], [15.0], [12.0], [8.0], [3.0]])
    np.testing.assert_allclose(nx.compute_value_targets(r, c, v, 1.0, 1.0)[:, 0], [3, 2, 15, 12, 8], atol=1e-5)
    np.testing.assert_allclose(
        nx.compute_value_targets(r, c, v, 0.997, 0.95)[:, 0],
        [2.994, 2.0, 14.933195, 11.967476, 7.991], rtol=1e-5)


def test_percentile_nearest():
    x = np.random.default_rng(0).standard_normal(15360).astype(np.float32)
    s = np.sort(x)
    assert nx.percentile_nearest(x, 5) == s[768]
    assert nx.percentile_nearest(x, 95) == s[14591]
    assert nx.percentile_nearest(np.arange(11.0), 50) == 5.0


def test_sample_categorical_contract():
    rng = np.random.default_rng(1)
    logits = rng.standard_normal((64, 32, 32)).astype(np.float32)
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p = (0.99 * p / p.sum(-1, keepdims=True) + 0.01 / 32).astype(np.float32)
    u = rng.random((64, 32), dtype=np.float32)
    idx = nx.sample_categorical(p, u)
    cdf = np.cumsum(p.astype(np.float64), -1)
    thr = u.astype(np.float64) * cdf[..., -1]
    take = np.take_along_axis(cdf, idx[..., None], -1)[..., 0]
    assert (take > thr).all()
    prev = np.where(idx > 0, np.take_along_axis(cdf, np.maximum(idx - 1, 0)[..., None], -1)[..., 0], 0.0)
    assert (prev <= thr).all()
    # u = 0 -> class 0; u -> 1 -> last class
    assert (nx.sample_categorical(p, np.zeros((64, 32), np.float32)) == 0).all()
    assert (nx.sample_categorical(p, np.full((64, 32), np.float32(1) - np.float32(2 ** -24))) == 31).all()
    # empirical frequencies follow probs
    pp = np.array([[0.7, 0.1, 0.1, 0.1]], np.float32).repeat(20000, 0)
    freq = np.bincount(nx.sample_categorical(pp, rng.random(20000, dtype=np.float32)), minlength=4) / 20000
    np.testing.assert_allclose(freq, [0.7, 0.1, 0.1, 0.1], atol=0.02)


def test_keras_adam_and_clip():
    g = [np.full((3,), 3.0, np.float32), np.full((4,), 4.0, np.float32)]
    clipped, gn = nx.clip_by_global_norm(g, 1.0)
    assert abs(gn - np.sqrt(27 + 64)) < 1e-6
    np.testing.assert_allclose(np.sqrt(sum((c ** 2).sum() for c in clipped)), 1.0, rtol=1e-6)
    same, _ = nx.clip_by_global_norm(g, 1000.0)
    np.testing.assert_array_equal(same[0], g[0])
    w, m, v = nx.keras_adam_step(np.ones(2, np.float32), np.array([0.5, -2.0], np.float32),
                                 np.zeros(2, np.float32), np.zeros(2, np.float32), 1, 1e-3, 1e-8)
    # first step of Adam moves by ~lr * sign(g)
    np.testing.assert_allclose(w, [1 - 1e-3, 1 + 1e-3], rtol=1e-5)
