"""Reference_Encoder / Style_Token_Layer / GST_Concated_Encoder with the reference's call
signatures (Modules/GST.py:12-124), executed by libgsttaco.so."""
from __future__ import annotations

from typing import Optional

from . import default_engine
from ..runtime import Engine


class _Layer:
    def __init__(self, engine: Optional[Engine] = None):
        self._engine = engine

    @property
    def engine(self) -> Engine:
        eng = self._engine or default_engine()
        if not eng.cfg.gst_use:
            # reference: Model.py:258-259 / 428-429
            raise NotImplementedError("GST is not used")
        return eng

    def __call__(self, inputs, **kw):
        return self.call(inputs, **kw)


class Reference_Encoder(_Layer):
    """Reference: Modules/GST.py:12-70.  inputs: [mels [B,T,mel], mel_lengths [B]] -> [B, Dense.Size]."""

    def call(self, inputs):
        mels, mel_lengths = inputs
        return self.engine.gst(mels, mel_lengths, drop_first=False, want=("ref",))["ref"]


class Style_Token_Layer(_Layer):
    """Reference: Modules/GST.py:72-109.  inputs: [mels_for_gst [B,1+T,mel] (initial frame is dropped,
    GST.py:98), mel_lengths [B]] -> [B, Attention.Size]."""

    def call(self, inputs, return_attention: bool = False):
        mels, mel_lengths = inputs
        want = ("gst", "attention") if return_attention else ("gst",)
        out = self.engine.gst(mels, mel_lengths, drop_first=True, want=want)
        return (out["gst"], out["attention"]) if return_attention else out["gst"]


class GST_Concated_Encoder(_Layer):
    """Reference: Modules/GST.py:111-124.  inputs: [encoders [B,T_v,D], gsts [B,S]] -> [B,T_v,S+D]
    (GST channels first).  Decoder callers can skip this layer altogether and pass
    ``enc_text=..., gst=...`` to Engine.decode: the concat is then folded into the value projection."""

    def call(self, inputs):
        encoders, gsts = inputs
        return self.engine.concat_encoder(encoders, gsts)
