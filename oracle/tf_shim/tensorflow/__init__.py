"""Minimal TensorFlow-2 API shim backed by torch (float64) - TEST INFRASTRUCTURE ONLY.

Purpose: let ``oracle/make_golden.py`` import and execute the reference's UNMODIFIED sources
(/root/reference/Modules/{Taco2,GST}.py, Modules/Attention/{Steps,Layers}.py) in a container where
TensorFlow cannot be installed, so that golden vectors come from the reference's own Python code:
decoder-step wiring, attention scoring, monotonic probability functions, safe_cumprod, multi-head
attention, Layer_Norm, loop conventions.  Only the Keras *primitives* (Dense, LSTMCell, GRU, Conv,
BatchNormalization, Dropout ...) are implemented here, independently of oracle/reference_port.py, from
the documented TF 2.x semantics.

Randomness is explicit: ``tf.random.normal`` and ``Dropout`` pop pre-seeded arrays from
``random._normal_queue`` / ``keras.layers.Dropout.mask_queue``.

Nothing in the product path or in the GPU tests imports this package.
"""
import builtins as _builtins
import math as _math
import types

import numpy as _np
import torch as _t

_t.set_default_dtype(_t.float64)

float32 = _t.float64  # everything is computed in float64 ("dtype" arguments are accepted and ignored)
float16 = _t.float64
float64 = _t.float64
int32 = _t.int32
int64 = _t.int64
bool_ = _t.bool


class TensorShape(object):
    def __init__(self, dims):
        self._dims = None if dims is None else [None if d is None else int(d) for d in dims]

    def as_list(self):
        return list(self._dims)

    def __getitem__(self, i):
        r = self._dims[i]
        return TensorShape(r) if isinstance(r, list) else r

    def __len__(self):
        return len(self._dims)

    def __iter__(self):
        return iter(self._dims)


_t.Tensor.get_shape = lambda self: TensorShape(list(self.shape))


def _T(x, dtype=None):
    if isinstance(x, _t.Tensor):
        return x
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], _t.Tensor):
        return _t.stack([_T(v) for v in x])
    a = _np.asarray(x)
    if a.dtype.kind == "f":
        return _t.as_tensor(a, dtype=_t.float64)
    return _t.as_tensor(a)


class _DType(object):
    """what `tensor.dtype` looks like to the reference: float32 (np.finfo(x.dtype.as_numpy_dtype).tiny, Steps.py:197)"""
    as_numpy_dtype = _np.float32


class _TT(_t.Tensor):
    @property
    def dtype(self):
        return _DType()


def convert_to_tensor(x, name=None, dtype=None):
    if isinstance(x, bool):
        return x
    x = _T(x)
    return x.as_subclass(_TT) if x.is_floating_point() else x


def constant(x, dtype=None):
    return _T(x)


def shape(x):
    return tuple(int(s) for s in x.shape)


def _shp(shape):
    if isinstance(shape, (int, _np.integer)):
        return [int(shape)]
    return [int(s) for s in shape]


def zeros(shape, dtype=None):
    return _t.zeros(_shp(shape), dtype=_t.float64)


def ones(shape, dtype=None):
    if dtype in (_t.int32, _t.int64):
        return _t.ones([int(s) for s in shape], dtype=dtype)
    return _t.ones([int(s) for s in shape], dtype=_t.float64)


def ones_like(x):
    return _t.ones_like(_T(x))


def concat(values, axis):
    return _t.cat([_T(v) for v in values], dim=axis)


def split(value, num_or_size_splits, axis=0):
    if isinstance(num_or_size_splits, int):
        return list(_t.chunk(value, num_or_size_splits, dim=axis))
    return list(_t.split(value, [int(s) for s in num_or_size_splits], dim=axis))


def stack(values, axis=0):
    return _t.stack([_T(v) for v in values], dim=axis)


def expand_dims(x, axis):
    return _T(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return x.squeeze() if axis is None else x.squeeze(axis)


def reshape(x, shape):
    return x.reshape([int(s) for s in shape])


def tile(x, multiples):
    return x.repeat([int(m) for m in multiples])


def _red(fn):
    def f(x, axis=None, keepdims=False):
        x = _T(x)
        if axis is None:
            return fn(x)
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        return fn(x, dim=ax, keepdim=keepdims)
    return f


reduce_sum = _red(_t.sum)
reduce_mean = _red(_t.mean)


def reduce_max(x, axis=None, keepdims=False):
    return x.max() if axis is None else x.amax(dim=axis, keepdim=keepdims)


def reduce_prod(x, axis=None):
    return int(_np.prod(_np.asarray(x)))


def tanh(x):
    return _t.tanh(x)


def sigmoid(x):
    return 1.0 / (1.0 + _t.exp(-x))


def exp(x):
    return _t.exp(x)


def square(x):
    return x * x


def maximum(a, b):
    return _t.maximum(_T(a), _T(b))


def clip_by_value(x, lo, hi):
    return _t.clamp(x, min=float(lo), max=float(hi))


def cumsum(x, axis=0, exclusive=False, reverse=False):
    assert not reverse
    c = _t.cumsum(x, dim=axis)
    return c - x if exclusive else c


def one_hot(indices, depth, dtype=None):
    out = _t.zeros(tuple(indices.shape) + (int(depth),), dtype=_t.float64)
    out.scatter_(-1, indices.long().unsqueeze(-1), 1.0)
    return out


def matmul(a, b, transpose_b=False):
    return a @ (b.transpose(-1, -2) if transpose_b else b)


def range(n):  # noqa: A001
    return _t.arange(int(n), dtype=_t.int32)


def gather_nd(params, indices):
    idx = indices.long()
    return params[tuple(idx[:, i] for i in _np.arange(idx.shape[1]))]


def cast(x, dtype):
    x = _T(x)
    if dtype in (_t.int32, _t.int64):
        return x.to(dtype)
    return x.to(_t.float64)


def less(a, b):
    return a < b


def greater_equal(a, b):
    return a >= b


def logical_not(x):
    return ~x


def logical_and(a, b):
    return a & b


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(pred) else false_fn()


def while_loop(cond, body, loop_vars, shape_invariants=None):
    v = list(loop_vars)
    while bool(cond(*v)):
        v = list(body(*v))
    return v


def zeros_initializer():
    return lambda shape: _t.zeros(shape, dtype=_t.float64)


def ones_initializer():
    return lambda shape: _t.ones(shape, dtype=_t.float64)


def _glorot(shape):
    shape = [int(s) for s in shape]
    fan_in = shape[0] if len(shape) >= 1 else 1
    fan_out = shape[-1] if len(shape) >= 1 else 1
    lim = _math.sqrt(6.0 / max(1, fan_in + fan_out))
    return (_t.rand(shape, dtype=_t.float64) * 2 - 1) * lim


def _init_from(spec):
    if callable(spec):
        return spec
    if spec in (None, "glorot_uniform"):
        return _glorot
    if spec == "zeros":
        return lambda shape: _t.zeros(shape, dtype=_t.float64)
    if spec == "ones":
        return lambda shape: _t.ones(shape, dtype=_t.float64)
    raise ValueError(spec)


math_ns = types.SimpleNamespace(
    log=_t.log, ceil=lambda x: _t.ceil(_T(x).to(_t.float64)), rsqrt=lambda x: 1.0 / _t.sqrt(x), tanh=_t.tanh, exp=_t.exp)
nest = types.SimpleNamespace()


def _map_structure(fn, s):
    if isinstance(s, (list, tuple)):
        return type(s)(_map_structure(fn, v) for v in s)
    return fn(s)


nest.map_structure = _map_structure


class _Random(object):
    def __init__(self):
        self._normal_queue = []

    def normal(self, shape, dtype=None, **kw):
        if not self._normal_queue:
            raise RuntimeError("tf.random.normal called with an empty explicit-noise queue")
        z = self._normal_queue.pop(0)
        assert tuple(z.shape) == tuple(int(s) for s in shape), (z.shape, shape)
        return z


random = _Random()


def _softmax(x, axis=-1):
    x = x - x.amax(dim=axis, keepdim=True)
    e = _t.exp(x)
    return e / e.sum(dim=axis, keepdim=True)


def _moments(x, axes, keepdims=False):
    ax = tuple(axes)
    m = x.mean(dim=ax, keepdim=True)
    v = ((x - m) ** 2).mean(dim=ax, keepdim=True)
    if not keepdims:
        m, v = m.squeeze(ax), v.squeeze(ax)
    return m, v


nn = types.SimpleNamespace(softmax=_softmax, tanh=_t.tanh, sigmoid=sigmoid, moments=_moments,
                           relu=lambda x: _t.clamp(x, min=0.0))


# --------------------------------------------------------------------------------------------
# Keras
# --------------------------------------------------------------------------------------------
class Layer(object):
    def __init__(self, *a, **kw):
        self.built = False
        self.dtype = float32
        self._weights = {}

    def add_weight(self, name=None, shape=None, initializer=None, dtype=None, trainable=True, **kw):
        w = _init_from(initializer)([int(s) for s in shape])
        self._weights[name] = w
        return w

    def build(self, input_shape):
        self.built = True

    def _maybe_build(self, inputs):
        if not getattr(self, "built", False):
            if isinstance(inputs, (list, tuple)):
                shp = [TensorShape(list(x.shape)) if isinstance(x, _t.Tensor) else None for x in inputs]
            else:
                shp = TensorShape(list(inputs.shape))
            self.build(shp)
            self.built = True

    def __call__(self, inputs=None, *args, **kwargs):
        if inputs is None and "inputs" in kwargs:
            inputs = kwargs.pop("inputs")
        self._maybe_build(inputs)
        # Keras injects `training` from the call context when the caller leaves it out
        import inspect
        params = inspect.signature(self.call).parameters
        if "training" not in kwargs and not args:
            if "training" in params and params["training"].default is inspect.Parameter.empty:
                kwargs["training"] = None
        elif "training" in kwargs and "training" not in params and not any(
                q.kind is inspect.Parameter.VAR_KEYWORD for q in params.values()):
            kwargs.pop("training")   # Keras drops the argument when call() does not take it (ConvBank, Highwaynet: Taco2.py:407,430)
        return self.call(inputs, *args, **kwargs)


class Model(Layer):
    pass


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, **kw):
        super(Dense, self).__init__()
        self.units, self.activation, self.use_bias = int(units), activation, use_bias

    def build(self, input_shape):
        self.kernel = _glorot([int(input_shape[-1]), self.units])
        self.bias = _t.zeros(self.units, dtype=_t.float64)

    def call(self, x, training=None):
        y = x @ self.kernel
        if self.use_bias:
            y = y + self.bias
        if self.activation in ("relu",):
            y = _t.clamp(y, min=0.0)
        elif self.activation in ("tanh",):
            y = _t.tanh(y)
        elif self.activation in ("sigmoid",):
            y = _t.sigmoid(y)
        elif callable(self.activation):
            y = self.activation(y)
        elif self.activation is not None:
            raise ValueError(self.activation)
        return y


class Dropout(Layer):
    mask_queue = []  # explicit keep masks, consumed in call order when training=True

    def __init__(self, rate, **kw):
        super(Dropout, self).__init__()
        self.rate = float(rate)

    def call(self, x, training=None):
        if not training or self.rate == 0.0:
            return x
        if not Dropout.mask_queue:
            raise RuntimeError("Dropout(training=True) called with an empty explicit-mask queue")
        keep = Dropout.mask_queue.pop(0)
        assert tuple(keep.shape) == tuple(x.shape), (keep.shape, x.shape)
        return x * keep / (1.0 - self.rate)


class ReLU(Layer):
    def call(self, x, training=None):
        return _t.clamp(x, min=0.0)


class Activation(Layer):
    def __init__(self, activation, **kw):
        super(Activation, self).__init__()
        self.activation = activation

    def call(self, x, training=None):
        return self.activation(x)


class Lambda(Layer):
    def __init__(self, function, **kw):
        super(Lambda, self).__init__()
        self.function = function

    def call(self, x, training=None):
        return self.function(x)


class BatchNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kw):
        super(BatchNormalization, self).__init__()
        self.epsilon = epsilon

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.gamma = _t.ones(c, dtype=_t.float64)
        self.beta = _t.zeros(c, dtype=_t.float64)
        self.moving_mean = _t.zeros(c, dtype=_t.float64)
        self.moving_variance = _t.ones(c, dtype=_t.float64)

    def call(self, x, training=None):
        if training:
            raise NotImplementedError("shim BatchNormalization implements the inference form only")
        return self.gamma * (x - self.moving_mean) / _t.sqrt(self.moving_variance + self.epsilon) + self.beta


def _same_pads(n, k, s):
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, **kw):
        super(Conv2D, self).__init__()
        self.filters, self.k, self.s, self.padding, self.use_bias = int(filters), int(kernel_size), int(strides), padding, use_bias

    def build(self, input_shape):
        self.kernel = _glorot([self.k, self.k, int(input_shape[-1]), self.filters])  # HWIO
        self.bias = _t.zeros(self.filters, dtype=_t.float64)

    def call(self, x, training=None):
        assert self.padding == "same"
        B, H, W, C = x.shape
        (pt, pb), (pl, pr) = _same_pads(H, self.k, self.s), _same_pads(W, self.k, self.s)
        xp = _t.zeros(B, H + pt + pb, W + pl + pr, C, dtype=_t.float64)
        xp[:, pt:pt + H, pl:pl + W] = x
        Ho, Wo = -(-H // self.s), -(-W // self.s)
        out = _t.zeros(B, Ho, Wo, self.filters, dtype=_t.float64)
        for kh in _np.arange(self.k):
            for kw_ in _np.arange(self.k):
                patch = xp[:, kh:kh + (Ho - 1) * self.s + 1:self.s, kw_:kw_ + (Wo - 1) * self.s + 1:self.s]
                out = out + patch @ self.kernel[kh, kw_]
        return out + self.bias if self.use_bias else out


class Conv1D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, **kw):
        super(Conv1D, self).__init__()
        self.filters, self.k, self.s, self.padding, self.use_bias = int(filters), int(kernel_size), int(strides), padding, use_bias

    def build(self, input_shape):
        self.kernel = _glorot([self.k, int(input_shape[-1]), self.filters])  # WIO
        self.bias = _t.zeros(self.filters, dtype=_t.float64)

    def call(self, x, training=None):
        assert self.padding == "same"
        B, W, C = x.shape
        pl, pr = _same_pads(W, self.k, self.s)
        xp = _t.zeros(B, W + pl + pr, C, dtype=_t.float64)
        xp[:, pl:pl + W] = x
        Wo = -(-W // self.s)
        out = _t.zeros(B, Wo, self.filters, dtype=_t.float64)
        for kw_ in _np.arange(self.k):
            out = out + xp[:, kw_:kw_ + (Wo - 1) * self.s + 1:self.s] @ self.kernel[kw_]
        return out + self.bias if self.use_bias else out


class LSTMCell(Layer):
    def __init__(self, units, recurrent_dropout=0.0, **kw):
        super(LSTMCell, self).__init__()
        assert recurrent_dropout == 0.0
        self.units = int(units)

    def build(self, input_shape):
        u = self.units
        self.kernel = _glorot([int(input_shape[-1]), 4 * u])
        self.recurrent_kernel = _glorot([u, 4 * u])
        self.bias = _t.zeros(4 * u, dtype=_t.float64)
        self.bias[u:2 * u] = 1.0  # unit_forget_bias

    def call(self, x, states):
        h, c = states
        u = self.units
        z = x @ self.kernel
        z = z + h @ self.recurrent_kernel
        z = z + self.bias
        i = sigmoid(z[:, 0 * u:1 * u])
        f = sigmoid(z[:, 1 * u:2 * u])
        g = _t.tanh(z[:, 2 * u:3 * u])
        o = sigmoid(z[:, 3 * u:4 * u])
        c2 = f * c + i * g
        h2 = o * _t.tanh(c2)
        return h2, [h2, c2]


class StackedRNNCells(Layer):
    def __init__(self, cells, **kw):
        super(StackedRNNCells, self).__init__()
        self.cells = list(cells)

    def get_initial_state(self, inputs=None, batch_size=None, dtype=None):
        return tuple([_t.zeros(int(batch_size), c.units, dtype=_t.float64), _t.zeros(int(batch_size), c.units, dtype=_t.float64)]
                     for c in self.cells)

    def __call__(self, inputs, states=None, **kw):
        x = inputs
        new_states = []
        for cell, st in zip(self.cells, states):
            cell._maybe_build(x)
            x, ns = cell.call(x, st)
            new_states.append(ns)
        return x, tuple(new_states)


class GRU(Layer):
    def __init__(self, units, return_sequences=False, **kw):
        super(GRU, self).__init__()
        self.units, self.return_sequences = int(units), return_sequences

    def build(self, input_shape):
        u = self.units
        self.kernel = _glorot([int(input_shape[-1]), 3 * u])
        self.recurrent_kernel = _glorot([u, 3 * u])
        self.bias = _t.zeros(2, 3 * u, dtype=_t.float64)  # reset_after=True

    def call(self, x, training=None):
        B, T, _ = x.shape
        u = self.units
        h = _t.zeros(B, u, dtype=_t.float64)
        outs = []
        for t in _np.arange(T):
            mx = x[:, t] @ self.kernel + self.bias[0]
            mh = h @ self.recurrent_kernel + self.bias[1]
            z = sigmoid(mx[:, :u] + mh[:, :u])
            r = sigmoid(mx[:, u:2 * u] + mh[:, u:2 * u])
            hh = _t.tanh(mx[:, 2 * u:] + r * mh[:, 2 * u:])
            h = z * h + (1.0 - z) * hh
            outs.append(h)
        seq = _t.stack(outs, dim=1)
        return seq if self.return_sequences else seq[:, -1]


class Sequential(Layer):
    def __init__(self, layers=None, **kw):
        super(Sequential, self).__init__()
        self.layers = list(layers or [])

    def add(self, layer):
        self.layers.append(layer)

    def __call__(self, inputs=None, training=None, **kw):
        x = inputs
        import inspect
        for l in self.layers:
            l._maybe_build(x)
            x = l.call(x, training=training) if "training" in inspect.signature(l.call).parameters else l.call(x)
        return x


class Embedding(Layer):
    """tf.keras.layers.Embedding (no mask): out = embeddings[ids]."""
    def __init__(self, input_dim, output_dim, **kw):
        super(Embedding, self).__init__()
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)

    def build(self, input_shape):
        self.embeddings = (_t.rand(self.input_dim, self.output_dim, dtype=_t.float64) - 0.5) * 0.1

    def call(self, x, training=None):
        return self.embeddings[_t.as_tensor(_np.asarray(x)).long()]


class LSTM(Layer):
    """tf.keras.layers.LSTM(units, return_sequences=True) over [B, T, C]: one LSTMCell driven step by step from a zero
    state (TF2 defaults: sigmoid recurrent activation, gate order i,f,c,o); ``go_backwards`` as Keras' Bidirectional uses it."""
    def __init__(self, units, recurrent_dropout=0.0, return_sequences=False, go_backwards=False, **kw):
        super(LSTM, self).__init__()
        assert return_sequences
        self.units = int(units)
        self.cell = LSTMCell(units, recurrent_dropout=recurrent_dropout)
        self.go_backwards = go_backwards

    def build(self, input_shape):
        self.cell.build(input_shape)

    def call(self, x, training=None):
        B, T = int(x.shape[0]), int(x.shape[1])
        h = _t.zeros(B, self.units, dtype=_t.float64)
        c = _t.zeros(B, self.units, dtype=_t.float64)
        outs = []
        order = list(_builtins.range(T))[::-1] if self.go_backwards else list(_builtins.range(T))
        for s in order:
            _, (h, c) = self.cell.call(x[:, s], [h, c])
            outs.append(h)
        return _t.stack(outs, dim=1)   # in processing order, like Keras


class Bidirectional(Layer):
    """tf.keras.layers.Bidirectional(layer, merge_mode='concat'): a forward copy and a go_backwards copy of ``layer``;
    the backward outputs are reversed in time before the concat."""
    def __init__(self, layer, **kw):
        super(Bidirectional, self).__init__()
        self.forward_layer = layer
        self.backward_layer = LSTM(layer.units, return_sequences=True, go_backwards=True)

    def build(self, input_shape):
        self.forward_layer.build(input_shape)
        self.backward_layer.build(input_shape)

    def call(self, x, training=None):
        f = self.forward_layer.call(x)
        b = _t.flip(self.backward_layer.call(x), dims=[1])
        return _t.cat([f, b], dim=-1)


class MaxPool1D(Layer):
    """tf.keras.layers.MaxPool1D on [B, W, C]: 'same' padding pads with -inf, pad_before = total // 2 (TF)."""

    def __init__(self, pool_size=2, strides=None, padding="valid", **kw):
        super(MaxPool1D, self).__init__()
        self.k, self.s, self.padding = int(pool_size), int(strides if strides is not None else pool_size), padding

    def call(self, x, training=None):
        B, W, C = x.shape
        pl, pr = _same_pads(W, self.k, self.s) if self.padding == "same" else (0, 0)
        xp = _t.full((B, W + pl + pr, C), float("-inf"), dtype=_t.float64)
        xp[:, pl:pl + W] = x
        Wo = (W + pl + pr - self.k) // self.s + 1
        out = xp[:, 0:(Wo - 1) * self.s + 1:self.s]
        for j in _np.arange(1, self.k):
            out = _t.maximum(out, xp[:, j:j + (Wo - 1) * self.s + 1:self.s])
        return out


class _NotOnHotPath(Layer):
    def __init__(self, *a, **kw):
        super(_NotOnHotPath, self).__init__()

    def call(self, *a, **kw):
        raise NotImplementedError("layer outside the hot path is not implemented by the shim")


class _BaseDenseAttention(Layer):
    """tf.keras.layers.Attention / AdditiveAttention skeleton (use_scale, causal, _validate_call_args)."""

    def __init__(self, use_scale=False, causal=False, **kw):
        super(_BaseDenseAttention, self).__init__()
        self.use_scale, self.causal = use_scale, causal
        self.scale = None

    def build(self, input_shape):
        self.scale = None  # use_scale=False everywhere on the hot path
        self.built = True

    def _validate_call_args(self, inputs, mask):
        if not isinstance(inputs, list):
            raise ValueError("inputs must be a list")
        if len(inputs) < 2 or len(inputs) > 3:
            raise ValueError("inputs must have 2 or 3 elements")


class _Schedule(object):
    def __init__(self, *a, **kw):
        pass


_layers = types.SimpleNamespace(
    Layer=Layer, Dense=Dense, Dropout=Dropout, ReLU=ReLU, Activation=Activation, Lambda=Lambda,
    BatchNormalization=BatchNormalization, Conv2D=Conv2D, Conv1D=Conv1D, LSTMCell=LSTMCell,
    StackedRNNCells=StackedRNNCells, GRU=GRU, Attention=_BaseDenseAttention, AdditiveAttention=_BaseDenseAttention,
    Embedding=Embedding, Bidirectional=Bidirectional, LSTM=LSTM, MaxPool1D=MaxPool1D,
    Input=lambda *a, **k: None)
_initializers = types.SimpleNamespace(
    TruncatedNormal=lambda stddev=0.05, **kw: (lambda shape: _t.clamp(_t.randn(shape, dtype=_t.float64) * stddev, -2 * stddev, 2 * stddev)),
    glorot_uniform=lambda: _glorot, constant=lambda v: (lambda shape: _t.zeros(shape, dtype=_t.float64) + float(_np.asarray(v).reshape(-1)[0])))
keras = types.SimpleNamespace(
    Model=Model, Sequential=Sequential, layers=_layers, initializers=_initializers,
    optimizers=types.SimpleNamespace(schedules=types.SimpleNamespace(ExponentialDecay=_Schedule, LearningRateSchedule=_Schedule)))
initializers = _initializers
math = math_ns
