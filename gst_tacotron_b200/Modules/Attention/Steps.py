"""BahdanauMonotonicAttention / StepwiseMonotonicAttention with the reference's constructor and call
signatures (Modules/Attention/Steps.py:51-229), executed by libgsttaco.so (gstk_attention_step).

Inside Decoder_Step these layers are fused into the persistent decoder kernel; the classes here are the
stand-alone drop-ins.  Variables (Steps.py:65-86) are created Keras-style at the first call (glorot-uniform
kernels, zero biases) and can be replaced with :meth:`set_weights`.  The sigmoid noise the reference draws
from TF's global RNG (Steps.py:220-221) is passed explicitly (``noise=``; None means no noise)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from .. import default_engine
from ... import _lib
from ...runtime import Engine, _ptr, _to_tensor


def _glorot(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[-1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


class BahdanauMonotonicAttention:
    _TYPE = "BMA"

    def __init__(self, size, sigmoid_noise=0.0, normalize=False, engine: Optional[Engine] = None, seed: int = 0, **kwargs):
        if normalize:
            raise ValueError("normalize=True (attention_g / attention_b, Steps.py:88-103) is not on the decode hot path")
        self.size = int(size)
        self.sigmoid_noise = float(sigmoid_noise)
        self.normalize = normalize
        self._engine = engine
        self._seed = seed
        self.weights = None

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    def build(self, query_dim, value_dim, key_dim=None):
        rng = np.random.default_rng(self._seed)
        w = {"Query/kernel": _glorot(rng, (query_dim, self.size)), "Query/bias": np.zeros(self.size, np.float32),
             "Value/kernel": _glorot(rng, (value_dim, self.size)), "Value/bias": np.zeros(self.size, np.float32),
             "attention_v": _glorot(rng, (self.size,)), "attention_score_bias": np.zeros((), np.float32)}
        if key_dim is not None:
            w["Key/kernel"] = _glorot(rng, (key_dim, self.size))
            w["Key/bias"] = np.zeros(self.size, np.float32)
        self.weights = w

    def set_weights(self, weights):
        self.weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}

    def initial_alignment_fn(self, batch_size, key_time, dtype=None):
        """Steps.py:201-206: one-hot at index 0."""
        a = torch.zeros(batch_size, key_time, device="cuda:{}".format(self.engine.device))
        a[:, 0] = 1.0
        return a

    def __call__(self, inputs, noise=None):
        return self.call(inputs, noise=noise)

    def call(self, inputs, noise=None):
        """inputs: [queries [B,Dq], values [B,T_v,Dv], previous_alignments [B,T_v]] or the 4-input form with
        keys [B,T_v,Dk] in third place (Steps.py:107-120) -> (context [B,size], alignment [B,T_v])."""
        if len(inputs) == 3:
            query, value, prev = inputs
            key = None
        elif len(inputs) == 4:
            query, value, key, prev = inputs
        else:
            raise ValueError("Unexpected input length")
        q, v, k, pa, nz = _to_tensor(query), _to_tensor(value), _to_tensor(key), _to_tensor(prev), _to_tensor(noise)
        if self.weights is None:
            self.build(int(q.shape[-1]), int(v.shape[-1]), None if k is None else int(k.shape[-1]))
        eng = self.engine
        B, Tv = int(v.shape[0]), int(v.shape[1])
        host = not (isinstance(v, torch.Tensor) and v.is_cuda)
        ctx, al = eng._alloc((B, self.size), host), eng._alloc((B, Tv), host)
        w = self.weights
        a = _lib.GstkAttentionArgs()
        a.batch, a.key_time, a.query_dim, a.value_dim, a.size = B, Tv, int(q.shape[-1]), int(v.shape[-1]), self.size
        a.key_dim = 0 if k is None else int(k.shape[-1])
        a.type = _lib.ATT[self._TYPE]
        a.sigmoid_noise = self.sigmoid_noise if nz is not None else 0.0
        a.query, a.value, a.key, a.prev_alignment, a.noise = _ptr(q), _ptr(v), _ptr(k), _ptr(pa), _ptr(nz)
        a.q_kernel, a.q_bias = w["Query/kernel"].ctypes.data, w["Query/bias"].ctypes.data
        a.v_kernel, a.v_bias = w["Value/kernel"].ctypes.data, w["Value/bias"].ctypes.data
        if k is not None:
            a.k_kernel, a.k_bias = w["Key/kernel"].ctypes.data, w["Key/bias"].ctypes.data
        sb = np.ascontiguousarray(w["attention_score_bias"], dtype=np.float32).reshape(1)
        a.attention_v, a.attention_score_bias = w["attention_v"].ctypes.data, sb.ctypes.data
        a.out_context, a.out_alignment = _ptr(ctx), _ptr(al)
        a.stream = eng._stream()
        eng._check(eng._lib.gstk_attention_step(eng._h, C.byref(a)))
        return ctx, al


class StepwiseMonotonicAttention(BahdanauMonotonicAttention):
    """Steps.py:208-229 (sigmoid_noise default 2.0)."""
    _TYPE = "SMA"

    def __init__(self, size, sigmoid_noise=2.0, normalize=False, **kwargs):
        super().__init__(size, sigmoid_noise, normalize, **kwargs)
