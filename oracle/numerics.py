"""ORACLE (test infrastructure, not product code) -- NumPy restatement of the small numeric
primitives on the DreamerV3 training hot path of sven1977/dreamer_v3.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module. The product path (`dreamer_v3_b200`) never does.

Parity pin: `two_hot` is pinned by the reference's own docstring known-answer vectors
(`utils/two_hot.py:23-37`, checked in tests/test_oracle_numerics.py). symlog / two_hot /
lambda-returns / percentile are additionally held to the outputs of the reference's own code run
over the TF-API shim (tests/golden/ref_*.npz, tests/test_reference_golden.py); TensorFlow's own
arithmetic of the primitives stays unpinned (TensorFlow is not installable in this image).
"""
import numpy as np

F32 = np.float32


def symlog(x, ft=F32):
    """`sign(x) * log(|x| + 1)` in float32 (reference utils/symlog.py:9-12).

    `|x| + 1` is a float32 add as in TF; the logarithm is taken in float64 and rounded
    once to float32 (i.e. the correctly rounded float32 log), which is the contract the
    CUDA path implements as well so that two-hot bucket indices are bit-exact.
    `ft=np.float64` runs the same formula in double (only for comparing against the float64 reference fixtures).
    """
    x = np.asarray(x).astype(ft)
    t = (np.abs(x) + ft(1.0)).astype(ft)
    return (np.sign(x) * np.log(t.astype(np.float64)).astype(ft)).astype(ft)


def inverse_symlog(y):
    """`sign(y) * (exp(|y|) - 1)` (reference utils/symlog.py:15-28)."""
    y = np.asarray(y).astype(F32)
    return (np.sign(y) * (np.exp(np.abs(y)) - F32(1.0))).astype(F32)


def two_hot_indices(value, num_buckets=255, lower_bound=-20.0, upper_bound=20.0, ft=F32):
    """(k, kp1, w_k, w_kp1) exactly as reference utils/two_hot.py:39-66 computes them,
    in float32 arithmetic (python-float constants are cast to float32 at use, as TF does)."""
    v = np.clip(np.asarray(value).astype(ft), ft(lower_bound), ft(upper_bound))
    delta = ft((upper_bound - lower_bound) / (num_buckets - 1))  # two_hot.py:44
    idx = ((ft(-lower_bound) + v).astype(ft) / delta).astype(ft)  # two_hot.py:46
    k = np.floor(idx)
    kp1 = np.ceil(idx)
    kp1 = np.where(k == kp1, kp1 + ft(1.0), kp1)  # two_hot.py:53
    kp1 = np.where(kp1 == ft(num_buckets), kp1 - ft(2.0), kp1)  # two_hot.py:58
    values_k = (ft(lower_bound) + (k * delta).astype(ft)).astype(ft)
    values_kp1 = (ft(lower_bound) + (kp1 * delta).astype(ft)).astype(ft)
    w_k = ((v - values_kp1).astype(ft) / (values_k - values_kp1).astype(ft)).astype(ft)
    w_kp1 = (ft(1.0) - w_k).astype(ft)
    return k.astype(np.int32), kp1.astype(np.int32), w_k, w_kp1


def two_hot(value, num_buckets=255, lower_bound=-20.0, upper_bound=20.0, ft=F32):
    """Dense two-hot encoding [n, num_buckets] (reference utils/two_hot.py:9-78)."""
    value = np.asarray(value).astype(ft).reshape(-1)
    k, kp1, w_k, w_kp1 = two_hot_indices(value, num_buckets, lower_bound, upper_bound, ft)
    out = np.zeros((value.shape[0], num_buckets), dtype=ft)
    rows = np.arange(value.shape[0])
    # tf.scatter_nd adds duplicates; k != kp1 always holds after the fix-ups above.
    np.add.at(out, (rows, k), w_k)
    np.add.at(out, (rows, kp1), w_kp1)
    return out


def compute_value_targets(rewards_t0_to_H, continues_t0_to_H, values_t0_to_H, gamma, lambda_,
                          intrinsic_rewards_t1_to_H=None):
    """Reverse lambda-return scan (reference losses/critic_loss.py:92-139).

    All inputs [H+1, N] in real (non-symlog) space; returns targets [H, N] for t0..H-1:
    R_H = v_H;  R_t = r_{t+1} + gamma c_{t+1} ((1-lambda) v_{t+1} + lambda R_{t+1}).
    """
    r = np.asarray(rewards_t0_to_H).astype(F32)[1:]
    if intrinsic_rewards_t1_to_H is not None:
        r = r + np.asarray(intrinsic_rewards_t1_to_H).astype(F32)
    v = np.asarray(values_t0_to_H).astype(F32)
    discount = (np.asarray(continues_t0_to_H).astype(F32)[1:] * F32(gamma)).astype(F32)
    inter = (r + discount * F32(1.0 - lambda_) * v[1:]).astype(F32)
    Rs = [v[-1]]
    for t in reversed(range(discount.shape[0])):
        Rs.append((inter[t] + discount[t] * F32(lambda_) * Rs[-1]).astype(F32))
    return np.stack(list(reversed(Rs))[:-1], axis=0)


def percentile_nearest(x, q):
    """tfp.stats.percentile(x, q) with its defaults (flatten, interpolation='nearest'),
    as used by reference losses/actor_loss.py:110-111. tfp sorts descending and gathers
    index round((n-1) * (1 - q/100)) (float64, round-half-even)."""
    flat = np.sort(np.asarray(x).astype(F32).reshape(-1))[::-1]
    n = flat.shape[0]
    idx = int(np.round((n - 1) * (1.0 - q / 100.0)))  # np.round is half-to-even
    return flat[idx]


def percentile_desc_index(n, q):
    return int(np.round((n - 1) * (1.0 - q / 100.0)))


def sample_categorical(probs, u):
    """Inverse-CDF categorical draw with explicit uniforms (the sampling contract of this
    repo): probs [..., K] float32 (unimixed), u [...] float32 in [0,1).
    The CDF is accumulated in float64 (exact for unimixed float32 probabilities, so the
    summation order does not matter) and the index is the first k with cdf[k] > u*total.
    Mirrors TF's CPU multinomial (double CDF + upper_bound(u * total))."""
    p = np.asarray(probs).astype(F32).astype(np.float64)
    cdf = np.cumsum(p, axis=-1)
    thr = np.asarray(u).astype(F32).astype(np.float64) * cdf[..., -1]
    idx = (cdf <= thr[..., None]).sum(axis=-1)
    return np.minimum(idx, p.shape[-1] - 1).astype(np.int32)


def cdf_margin(probs, u):
    """Smallest |cdf[k]/total - u| over k: how close a uniform is to flipping the drawn index."""
    p = np.asarray(probs).astype(F32).astype(np.float64)
    cdf = np.cumsum(p, axis=-1)
    cdf = cdf / cdf[..., -1:]
    return np.abs(cdf[..., :-1] - np.asarray(u).astype(np.float64)[..., None]).min(axis=-1)


def clip_by_global_norm(grads, clip):
    """tf.clip_by_global_norm: g * clip / max(||g||_2, clip) (train_one_step.py:82)."""
    gn = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    scale = clip / max(gn, clip)
    return [(g * F32(scale)).astype(F32) for g in grads], gn


def keras_adam_step(w, g, m, v, t, lr, eps, beta1=0.9, beta2=0.999):
    """One Keras-2 Adam update: epsilon is added to the *uncorrected* sqrt(v).
    `t` is the 1-based step count. Returns (w, m, v)."""
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = m + (g - m) * F32(1.0 - beta1)
    v = v + (g * g - v) * F32(1.0 - beta2)
    w = w - F32(lr_t) * m / (np.sqrt(v) + F32(eps))
    return w.astype(F32), m.astype(F32), v.astype(F32)
