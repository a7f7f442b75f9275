"""tensorflow_probability subset used by the reference's hot path, on the PyTorch-backed `tensorflow` shim.
TEST INFRASTRUCTURE ONLY.  Formulas follow tfp's definitions:

* OneHotCategorical(logits | probs): `logits` is returned as passed; entropy / KL use log_softmax of it (tfp
  re-normalises); `probs=` means logits = log(probs).
* Independent(d, n): sums log_prob / entropy / KL over the last n batch axes.
* Bernoulli(logits): mode = 1[logits > 0]; log_prob(x) = x log sigmoid(l) + (1 - x) log sigmoid(-l).
* Normal(loc, scale): reparameterised sample loc + scale * eps.
* MultivariateNormalDiag(loc, scale_diag): `.loc`, `.sample()`.
* stats.percentile(x, q): flattened, interpolation "nearest": sorted[round((n - 1) q / 100)] (float64, half-even).

Every `sample()` takes its randomness from the generator `RNG` installed by the calling script and logs the numbers
it used in `DRAW_LOG` in call order ("uniform" for categorical draws - index = first k with cdf[k] > u, CDF of the
float32 probabilities accumulated in float64 - and "normal" for Gaussian draws); the script saves them with the
outputs so that the oracle and the CUDA engine can be fed exactly the same noise.
"""
import types

import numpy as np
import torch

import tensorflow as tf

RNG = None          # numpy Generator the calling script installs; the only source of randomness
MARGIN = 1e-3       # categorical uniforms are redrawn until no CDF boundary is closer than this
DRAW_LOG = []       # (kind, ndarray actually used) of every draw, in call order - the script saves these


def _uniform(p):
    """Uniforms for a categorical draw over probabilities p[..., K] (float32 values, float64 CDF), redrawn where a
    CDF boundary is within MARGIN so that a last-ulp difference in p cannot change the index."""
    cdf = np.cumsum(p.astype(np.float64), axis=-1)
    u = RNG.random(p.shape[:-1], dtype=np.float32)
    for _ in range(200):
        bad = np.abs(cdf - u[..., None].astype(np.float64)).min(-1) < MARGIN
        if not bad.any():
            break
        u[bad] = RNG.random(int(bad.sum()), dtype=np.float32)
    DRAW_LOG.append(("uniform", u.copy()))
    return u, cdf


def _normal(shape):
    eps = RNG.standard_normal(tuple(int(s) for s in shape)).astype(np.float32)
    DRAW_LOG.append(("normal", eps.copy()))
    return eps


def _wrap(t):
    return t.as_subclass(tf.Tensor)


class OneHotCategorical:
    def __init__(self, logits=None, probs=None, dtype=None):
        assert (logits is None) != (probs is None)
        self._probs_given = probs
        self.logits = logits if logits is not None else _wrap(torch.log(probs))
        self.dtype = dtype or tf.int32

    def _log_p(self):
        return torch.log_softmax(self.logits, dim=-1)

    def sample(self):
        p = torch.softmax(self.logits.detach().as_subclass(torch.Tensor), dim=-1)
        u, cdf = _uniform(p.to(torch.float32).numpy())
        idx = np.minimum((cdf <= u[..., None].astype(np.float64)).sum(-1), p.shape[-1] - 1)
        return _wrap(torch.nn.functional.one_hot(torch.as_tensor(idx, dtype=torch.int64), p.shape[-1]).to(self.dtype))

    def log_prob(self, x):
        return _wrap((x * self._log_p()).sum(-1))

    def entropy(self):
        lp = self._log_p()
        return _wrap(-(torch.exp(lp) * lp).sum(-1))

    def _kl(self, other):
        lp, lq = self._log_p(), other._log_p()
        return (torch.exp(lp) * (lp - lq)).sum(-1)


class Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def sample(self):
        eps = _normal(self.loc.shape)
        return _wrap(self.loc + self.scale * torch.as_tensor(eps, dtype=self.loc.dtype))

    def log_prob(self, x):
        return _wrap(-0.5 * ((x - self.loc) / self.scale) ** 2 - torch.log(self.scale) - 0.5 * np.log(2.0 * np.pi))

    def entropy(self):
        return _wrap(0.5 + 0.5 * np.log(2.0 * np.pi) + torch.log(self.scale))


class Independent:
    def __init__(self, distribution, reinterpreted_batch_ndims=None):
        self.distribution = distribution
        self.n = int(reinterpreted_batch_ndims)

    def _sum(self, x):
        return _wrap(x.sum(dim=tuple(range(-self.n, 0)))) if self.n else _wrap(x)

    def sample(self):
        return self.distribution.sample()

    def log_prob(self, x):
        return self._sum(self.distribution.log_prob(x))

    def entropy(self):
        return self._sum(self.distribution.entropy())

    def _kl(self, other):
        return self._sum(self.distribution._kl(other.distribution))


class Bernoulli:
    def __init__(self, logits=None, dtype=None):
        self.logits = logits
        self.dtype = dtype or tf.int32

    def mode(self):
        return _wrap((self.logits > 0).to(self.dtype))

    def log_prob(self, x):
        ls = torch.nn.functional.logsigmoid
        return _wrap(x * ls(self.logits) + (1.0 - x) * ls(-self.logits))


class MultivariateNormalDiag:
    def __init__(self, loc=None, scale_diag=None):
        self.loc, self.scale_diag = loc, scale_diag

    def sample(self):
        eps = _normal(self.loc.shape)
        return _wrap(self.loc + self.scale_diag * torch.as_tensor(eps, dtype=self.loc.dtype))


def kl_divergence(a, b):
    return _wrap(a._kl(b))


def percentile(x, q, **kw):
    assert not kw, "only the defaults (all axes, interpolation='nearest') are restated"
    flat = torch.sort(x.detach().as_subclass(torch.Tensor).reshape(-1)).values
    n = flat.numel()
    return _wrap(flat[int(np.round((n - 1) * (np.float64(q) / 100.0)))])


distributions = types.SimpleNamespace(
    OneHotCategorical=OneHotCategorical, Normal=Normal, Independent=Independent, Bernoulli=Bernoulli,
    MultivariateNormalDiag=MultivariateNormalDiag, kl_divergence=kl_divergence)
stats = types.SimpleNamespace(percentile=percentile)
