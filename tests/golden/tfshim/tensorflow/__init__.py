"""A *minimal* stand-in for the `tensorflow` API surface that sven1977/dreamer_v3's hot path uses, backed by
PyTorch-CPU tensors and autograd.  TEST INFRASTRUCTURE ONLY (tests/golden/make_reference_golden.py).

Why: TensorFlow / Keras / tfp are not installable in the build image, so the reference cannot run as-is.  With this
shim on `sys.path` the reference's OWN Python files under /root/reference (models/, losses/, training/, utils/) are
imported and executed unmodified: the composition - masking, concatenation orders, loop structure, stop-gradients,
loss weighting, reductions, optimizer calls - is the reference's code; only the ~60 primitive ops below and the Keras
layers / tfp distributions are restated here (SURVEY.md Appendix A.1 lists the Keras / tfp defaults they follow).

`tf.float32` maps to `FLOAT` (torch.float32, or torch.float64 after `set_float(torch.float64)`), so the same
reference code can be run in double precision to produce tight golden vectors.

Random draws never come from a generator: every `sample()` of a tfp distribution and every initializer takes its
numbers from the generator the calling script installs (`tensorflow_probability.RNG`); the numbers used are logged.
"""
import builtins
import sys
import types

import numpy as np
import torch

FLOAT = torch.float32


def set_float(dtype):
    global FLOAT, float32
    FLOAT = dtype
    float32 = dtype
    sys.modules[__name__].float32 = dtype


float32 = FLOAT
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
uint8 = torch.uint8
bool = torch.bool  # noqa: A001  (tf.bool)


class TensorShape(tuple):
    def as_list(self):
        return list(self)

    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return TensorShape(r) if isinstance(i, slice) else r


class Tensor(torch.Tensor):
    """torch.Tensor with the handful of tf.Tensor members the reference touches."""

    @property
    def shape(self):
        return TensorShape(torch.Tensor.shape.__get__(self))

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()

    def __bool__(self):
        return builtins.bool(self.detach().as_subclass(torch.Tensor).item())

    # tf.Tensor is immutable: `a += b` rebinds the name to a new tensor. torch's in-place versions would also change every
    # view of `a` (critic_loss.py:117-120 adds the intrinsic rewards to a slice of dream_data["rewards_dreamed_t0_to_H_B"]).
    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other


def _wrap(t):
    return t.as_subclass(Tensor) if not isinstance(t, Tensor) else t


def _dt(dtype):
    return FLOAT if dtype is None else dtype


def convert_to_tensor(value, dtype=None, **_):
    if isinstance(value, torch.Tensor):
        t = value if dtype is None else value.to(dtype)
    else:
        a = np.asarray(value)
        if dtype is None and a.dtype.kind == "f":
            dtype = FLOAT  # python floats are float32 in TF; the golden script passes arrays in FLOAT already
        t = torch.as_tensor(a, dtype=dtype)
    return _wrap(t)


def is_tensor(x):
    return isinstance(x, torch.Tensor)


class Variable(Tensor):
    """tf.Variable: a leaf tensor; `trainable` decides whether autograd tracks it."""

    def __new__(cls, initial_value, trainable=True, dtype=None, name=None):
        if isinstance(initial_value, torch.Tensor):
            data = initial_value.detach().clone().to(_dt(dtype) if dtype is not None else initial_value.dtype)
        else:
            data = torch.as_tensor(np.asarray(initial_value), dtype=_dt(dtype))
        v = data.as_subclass(cls)
        v.requires_grad_(builtins.bool(trainable) and data.dtype.is_floating_point)
        v.trainable = builtins.bool(trainable)
        v.var_name = name
        return v

    def assign(self, value):
        with torch.no_grad():
            self.copy_(torch.as_tensor(value, dtype=self.dtype))
        return self

    def assign_sub(self, value):
        with torch.no_grad():
            self.sub_(torch.as_tensor(value, dtype=self.dtype))
        return self


def function(fn=None, **_):
    """tf.function: eager execution is the only mode here."""
    if fn is None:
        return lambda f: f
    return fn


# --------------------------------------------------------------------------------------- shapes
def shape(x):
    return list(x.shape)


def _ints(shape_):
    if isinstance(shape_, torch.Tensor):
        return [int(s) for s in shape_.tolist()] if shape_.dim() else [int(shape_)]
    if isinstance(shape_, (int, np.integer)):
        return [int(shape_)]
    return [int(s) for s in shape_]


def reshape(x, shape, name=None):  # noqa: A002
    return _wrap(torch.reshape(convert_to_tensor(x), _ints(shape)))


def concat(values, axis, name=None):
    if not any(isinstance(v, torch.Tensor) for v in values):  # shape arithmetic on python lists
        out = []
        for v in values:
            out += list(v)
        return out
    return _wrap(torch.cat([convert_to_tensor(v) for v in values], dim=axis))


def stack(values, axis=0, name=None):
    return _wrap(torch.stack([convert_to_tensor(v) for v in values], dim=axis))


def transpose(x, perm=None):
    return _wrap(x.permute(*perm)) if perm is not None else _wrap(x.T)


def expand_dims(x, axis):
    return _wrap(torch.unsqueeze(convert_to_tensor(x), axis))


def squeeze(x, axis=None):
    return _wrap(torch.squeeze(x) if axis is None else torch.squeeze(x, axis))


def repeat(x, repeats, axis=None):
    return _wrap(torch.repeat_interleave(convert_to_tensor(x), int(repeats), dim=axis))


def zeros(shape, dtype=None):  # noqa: A002
    return _wrap(torch.zeros(_ints(shape), dtype=_dt(dtype)))


def ones_like(x):
    return _wrap(torch.ones_like(x))


def fill(dims, value):
    return _wrap(torch.full(_ints(dims), value, dtype=FLOAT))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _wrap(torch.arange(start, limit, delta, dtype=dtype))


def linspace(start, stop, num):
    return _wrap(torch.linspace(float(start), float(stop), int(num), dtype=FLOAT))


def cast(x, dtype):
    return _wrap(convert_to_tensor(x).to(dtype))


def one_hot(indices, depth, dtype=None):
    idx = convert_to_tensor(indices).to(torch.int64)
    return _wrap(torch.nn.functional.one_hot(idx, int(depth)).to(_dt(dtype)))


def argmax(x, axis=None):
    return _wrap(torch.argmax(x, dim=axis))


def where(cond, x, y):
    return _wrap(torch.where(cond, x, y))


def stop_gradient(x):
    return _wrap(x.detach())


def einsum(eq, *ops):
    return _wrap(torch.einsum(eq, *ops))


def multiply(a, b):
    return _wrap(a * b)


def clip_by_value(x, lo, hi):
    return _wrap(torch.clamp(x, float(lo), float(hi)))


def scatter_nd(indices, updates, shape):  # noqa: A002
    """Sums `updates` into a zero tensor at `indices` (duplicates accumulate), rank-2 indices only."""
    out = torch.zeros(_ints(shape), dtype=updates.dtype)
    idx = indices.to(torch.int64)
    out = out.index_put(tuple(idx[:, i] for i in builtins.range(idx.shape[1])), updates, accumulate=True)
    return _wrap(out)


def tanh(x):
    return _wrap(torch.tanh(x))


def sigmoid(x):
    return _wrap(torch.sigmoid(x))


def _reduce(fn):
    def red(x, axis=None, keepdims=False, name=None):
        if isinstance(x, (list, tuple)):
            x = torch.stack([convert_to_tensor(v) for v in x])
        if axis is None:
            return _wrap(fn(x))
        return _wrap(fn(x, dim=axis, keepdim=keepdims))
    return red


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)


def _reduce_max(x, axis=None, keepdims=False):
    if isinstance(x, (list, tuple)):
        x = torch.stack([convert_to_tensor(v) for v in x])
    return _wrap(torch.max(x)) if axis is None else _wrap(torch.amax(x, dim=axis, keepdim=keepdims))


reduce_max = _reduce_max


def assert_equal(a, b, message=None):
    if not builtins.bool(torch.all(torch.as_tensor(a) == torch.as_tensor(b))):
        raise AssertionError(message or f"tf.assert_equal failed: {a} != {b}")


def clip_by_global_norm(t_list, clip_norm):
    """g * clip / max(||g||, clip) with ||g|| over every tensor of the list."""
    t_list = list(t_list)
    sq = sum(torch.sum(g.detach() ** 2) for g in t_list if g is not None)
    norm = torch.sqrt(sq)
    clip = torch.as_tensor(clip_norm, dtype=norm.dtype)
    scale = clip / torch.maximum(norm, clip)
    return [None if g is None else _wrap(g * scale) for g in t_list], _wrap(norm)


class GradientTape:
    """Autograd records everywhere in PyTorch; the tape only supplies `.gradient`."""

    def __init__(self, persistent=False, **_):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, target, sources):
        sources = list(sources)
        grads = torch.autograd.grad(target, sources, retain_graph=True, allow_unused=True)
        return [None if g is None else _wrap(g) for g in grads]


# --------------------------------------------------------------------------------------- tf.math / tf.nn / ...
math = types.ModuleType("tensorflow.math")
math.abs = lambda x: _wrap(torch.abs(convert_to_tensor(x)))
math.sign = lambda x: _wrap(torch.sign(convert_to_tensor(x)))
math.log = lambda x: _wrap(torch.log(convert_to_tensor(x)))
math.exp = lambda x: _wrap(torch.exp(convert_to_tensor(x)))
math.square = lambda x: _wrap(torch.square(x))
math.floor = lambda x: _wrap(torch.floor(x))
math.ceil = lambda x: _wrap(torch.ceil(x))
math.tanh = tanh
math.is_nan = lambda x: _wrap(torch.isnan(x))
math.cumprod = lambda x, axis=0: _wrap(torch.cumprod(x, dim=axis))
math.reduce_std = lambda x, axis=None: _wrap(torch.std(x, unbiased=False) if axis is None
                                             else torch.std(x, dim=axis, unbiased=False))


def _maximum(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    dt = a.dtype if a.dim() else b.dtype
    return _wrap(torch.maximum(a.to(dt), b.to(dt)))


math.maximum = _maximum
math.reduce_logsumexp = lambda x, axis=None, keepdims=False: _wrap(torch.logsumexp(x, dim=axis, keepdim=keepdims))

nn = types.ModuleType("tensorflow.nn")
nn.silu = lambda x: _wrap(torch.nn.functional.silu(x))
nn.softmax = lambda x, axis=-1: _wrap(torch.softmax(x, dim=axis))

random = types.ModuleType("tensorflow.random")
config = types.ModuleType("tensorflow.config")
config.list_physical_devices = lambda *_a, **_k: []
config.experimental = types.SimpleNamespace(set_memory_growth=lambda *a, **k: None,
                                            get_memory_info=lambda *a, **k: {"current": 0})
losses = types.ModuleType("tensorflow.losses")
losses.mse = lambda a, b: _wrap(torch.mean((a - b) ** 2, dim=-1))

from . import keras  # noqa: E402

for _m in (math, nn, random, config, losses):
    sys.modules[_m.__name__] = _m
