"""tf.keras subset: Layer / Model bookkeeping, Dense, LayerNormalization, GRU (single step), Conv2D,
Conv2DTranspose, Flatten, Adam.  Semantics restated from the Keras-2 defaults; variable
values are never drawn here - layers are built with zeros and the golden script assigns every weight by attribute
path afterwards."""
import sys
import types

import torch

import tensorflow as tf


def _wrap(t):
    return t.as_subclass(tf.Tensor) if not isinstance(t, tf.Tensor) else t


class Layer:
    def __init__(self, trainable=True, name=None, **_):
        self.trainable = trainable
        self.name = name
        self.built = False

    # -- variable tracking: walk attributes in definition order -------------------------------------------------
    def _walk(self, seen):
        for v in self.__dict__.values():
            yield from _walk_value(v, seen)

    @property
    def variables(self):
        return list(self._walk(set()))

    @property
    def trainable_variables(self):
        return [v for v in self.variables if v.trainable]

    def add_weight(self, shape, trainable=True):
        return tf.Variable(torch.zeros(tuple(int(s) for s in shape), dtype=tf.FLOAT),
                           trainable=trainable and self.trainable)

    def build(self, input_shape):
        pass

    def __call__(self, *args, **kwargs):
        if not self.built:
            first = args[0] if args else next(iter(kwargs.values()))
            if isinstance(first, torch.Tensor):
                self.build(tuple(first.shape))
            self.built = True
        return self.call(*args, **kwargs)


def _walk_value(v, seen):
    if isinstance(v, tf.Variable):
        if id(v) not in seen:
            seen.add(id(v))
            yield v
    elif isinstance(v, Layer):
        if id(v) not in seen:
            seen.add(id(v))
            yield from v._walk(seen)
    elif isinstance(v, (list, tuple)):
        for e in v:
            yield from _walk_value(e, seen)
    elif isinstance(v, dict):
        for e in v.values():
            yield from _walk_value(e, seen)


class Model(Layer):
    def summary(self, *a, **k):
        pass


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
                 trainable=True, name=None):
        super().__init__(trainable=trainable, name=name)
        assert activation is None
        self.units, self.use_bias = int(units), use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight((input_shape[-1], self.units))
        self.bias = self.add_weight((self.units,)) if self.use_bias else None

    def call(self, x):
        y = tf.convert_to_tensor(x) @ self.kernel
        return _wrap(y + self.bias if self.use_bias else y)


class LayerNormalization(Layer):
    """Keras default: axis=-1, epsilon=1e-3, biased variance, gamma/beta of the last axis' width."""

    def __init__(self, trainable=True, epsilon=1e-3, name=None):
        super().__init__(trainable=trainable, name=name)
        self.epsilon = epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight((input_shape[-1],))
        with torch.no_grad():
            self.gamma.fill_(1.0)
        self.beta = self.add_weight((input_shape[-1],))

    def call(self, inputs):
        x = inputs
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
        return _wrap((x - mean) * torch.rsqrt(var + self.epsilon) * self.gamma + self.beta)


class GRU(Layer):
    """Keras GRU (TF2 defaults: tanh / sigmoid, use_bias, reset_after=True, gate order z, r, h), time-major input
    [T, B, D] with `initial_state`; returns the last output."""

    def __init__(self, units, return_sequences=False, return_state=False, time_major=False, trainable=True):
        super().__init__(trainable=trainable)
        assert time_major and not return_sequences and not return_state
        self.units = int(units)

    def build(self, input_shape):
        G = self.units
        self.kernel = self.add_weight((input_shape[-1], 3 * G))
        self.recurrent_kernel = self.add_weight((G, 3 * G))
        self.bias = self.add_weight((2, 3 * G))

    def call(self, inputs, initial_state=None):
        G = self.units
        h = initial_state
        for t in range(inputs.shape[0]):
            gx = inputs[t] @ self.kernel + self.bias[0]
            gh = h @ self.recurrent_kernel + self.bias[1]
            z = torch.sigmoid(gx[:, :G] + gh[:, :G])
            r = torch.sigmoid(gx[:, G:2 * G] + gh[:, G:2 * G])
            c = torch.tanh(gx[:, 2 * G:] + r * gh[:, 2 * G:])
            h = z * h + (1.0 - z) * c
        return _wrap(h)


class Conv2D(Layer):
    """NHWC, kernel [kh, kw, cin, cout], padding "same" with k=4, s=2 on even sizes = one pixel each side."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", activation=None, use_bias=True,
                 trainable=True):
        super().__init__(trainable=trainable)
        assert kernel_size == 4 and tuple(strides) == (2, 2) and padding == "same" and activation is None
        self.filters, self.use_bias = int(filters), use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight((4, 4, input_shape[-1], self.filters))
        self.bias = self.add_weight((self.filters,)) if self.use_bias else None

    def call(self, inputs):
        x = inputs.permute(0, 3, 1, 2)
        y = torch.nn.functional.conv2d(x, self.kernel.permute(3, 2, 0, 1), self.bias, stride=2, padding=1)
        return _wrap(y.permute(0, 2, 3, 1))


class Conv2DTranspose(Layer):
    """NHWC, kernel [kh, kw, cout, cin]; "same" with k=4, s=2 doubles the size (the adjoint of the Conv2D above)."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", activation=None, use_bias=True,
                 trainable=True):
        super().__init__(trainable=trainable)
        assert kernel_size == 4 and tuple(strides) == (2, 2) and padding == "same" and activation is None
        self.filters, self.use_bias = int(filters), use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight((4, 4, self.filters, input_shape[-1]))
        self.bias = self.add_weight((self.filters,)) if self.use_bias else None

    def call(self, inputs):
        x = inputs.permute(0, 3, 1, 2)
        y = torch.nn.functional.conv_transpose2d(x, self.kernel.permute(3, 2, 0, 1), self.bias, stride=2, padding=1)
        return _wrap(y.permute(0, 2, 3, 1))


class Flatten(Layer):
    def __init__(self, data_format=None):
        super().__init__()

    def call(self, x):
        return _wrap(x.reshape(x.shape[0], -1))


class Adam:
    """Keras-2 Adam: w -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps) (eps outside the correction)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self.m, self.v = {}, {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        alpha = self.lr * (1.0 - self.b2 ** t) ** 0.5 / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for g, var in grads_and_vars:
                if g is None:
                    continue
                g = g.detach().as_subclass(torch.Tensor)
                m = self.m.setdefault(id(var), torch.zeros_like(g))
                v = self.v.setdefault(id(var), torch.zeros_like(g))
                m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                var.as_subclass(torch.Tensor).sub_(alpha * m / (torch.sqrt(v) + self.eps))


layers = types.ModuleType("tensorflow.keras.layers")
for _c in (Layer, Dense, LayerNormalization, GRU, Conv2D, Conv2DTranspose, Flatten):
    setattr(layers, _c.__name__, _c)
optimizers = types.ModuleType("tensorflow.keras.optimizers")
optimizers.Adam = Adam
models = types.ModuleType("tensorflow.keras.models")
for _m in (layers, optimizers, models):
    sys.modules[_m.__name__] = _m
