"""MultiHeadAttention with the reference's constructor and call signature
(Modules/Attention/Layers.py:147-228; residual + Layer_Norm 254-285), executed by libgsttaco.so (gstk_mha).

Inside Style_Token_Layer the attention is fused with the reference encoder's GRU/Dense epilogue; this class is
the stand-alone drop-in for arbitrary [query, value] inputs."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import default_engine
from ...runtime import Engine


class MultiHeadAttention:
    def __init__(self, num_heads, size, use_scale=False, engine: Optional[Engine] = None, seed: int = 0, **kwargs):
        if size % num_heads != 0:
            # same message as the reference, Layers.py:155-156
            raise ValueError("size must be divisible by num_heads. ('{}' % '{}' != 0)".format(size, num_heads))
        if use_scale:
            raise ValueError("use_scale=True is not on the decode hot path (reference default False, Layers.py:152)")
        self.num_heads, self.size, self.use_scale = int(num_heads), int(size), use_scale
        self._engine, self._seed = engine, seed
        self.weights = None

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    def build(self, query_dim, value_dim):
        rng = np.random.default_rng(self._seed)

        def glorot(i, o):
            lim = np.sqrt(6.0 / (i + o))
            return rng.uniform(-lim, lim, size=(i, o)).astype(np.float32)
        self.weights = {"Query/kernel": glorot(query_dim, self.size), "Query/bias": np.zeros(self.size, np.float32),
                        "Value/kernel": glorot(value_dim, self.size), "Value/bias": np.zeros(self.size, np.float32),
                        "Layer_Normalization/gamma": np.ones(self.size, np.float32),
                        "Layer_Normalization/beta": np.zeros(self.size, np.float32)}

    def set_weights(self, weights):
        self.weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}

    def __call__(self, inputs, mask=None):
        return self.call(inputs, mask=mask)

    def call(self, inputs, mask=None):
        """inputs: [query [B,Tq,Dq], value [B,Tv,Dv]] (key = value) -> (result [B,Tq,size], distribution [B,Tq,Tv])."""
        if not isinstance(inputs, (list, tuple)) or len(inputs) != 2:
            raise ValueError("MultiHeadAttention drop-in supports the [query, value] form the reference uses (GST.py:104-107)")
        if mask is not None:
            raise ValueError("masks are not used on the decode hot path")
        q, v = inputs
        if self.weights is None:
            self.build(int(np.shape(q)[-1]), int(np.shape(v)[-1]))
        w = self.weights
        return self.engine.mha(q, v, w["Query/kernel"], w["Query/bias"], w["Value/kernel"], w["Value/bias"],
                               w["Layer_Normalization/gamma"], w["Layer_Normalization/beta"], self.num_heads)
