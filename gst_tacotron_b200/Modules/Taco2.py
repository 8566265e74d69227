"""Decoder_Step / Decoder with the reference's call signatures (Modules/Taco2.py:53-232), executed
by libgsttaco.so.

Randomness: the reference draws prenet dropout masks (Taco2.py:283, always on) and SMA sigmoid
noise (Steps.py:220-221) from TF's global RNG.  The drop-in exposes that explicitly through
keyword arguments that default to the library's counter-based Philox stream:
``rng='philox'|'external'|'none'``, ``seed``, and for 'external' ``keep0/keep1/noise`` tensors."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import default_engine
from ..runtime import Engine


class Encoder:
    """Reference: Modules/Taco2.py:12-51.  ``Encoder()(tokens, training)`` -> [B, T_v, 2 * RNN.Size];
    the conv stack and the LSTM input projections run on the implicit-GEMM kernels, the recurrence in one kernel per call.
    Like the reference at inference, dropout is off and there is no padding mask."""

    def __init__(self, engine: Optional[Engine] = None):
        self._engine = engine

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    def __call__(self, inputs, training=False):
        return self.call(inputs, training)

    def call(self, inputs, training=False):
        if training:
            raise NotImplementedError("the Encoder drop-in is inference-only (conv Dropout is not applied)")
        return self.engine.encoder(inputs)


class Decoder_Step:
    """Reference: Modules/Taco2.py:53-120."""

    def __init__(self, engine: Optional[Engine] = None):
        self._engine = engine
        self._step = 0

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    # -- reference: Taco2.py:91 (StackedRNNCells.get_initial_state)
    def get_initial_state(self, inputs=None, batch_size=None, dtype=None):
        cfg = self.engine.cfg
        dev = "cuda:{}".format(self.engine.device)
        return tuple((torch.zeros(batch_size, u, device=dev), torch.zeros(batch_size, u, device=dev))
                     for u in cfg.lstm_sizes)

    # -- reference: Taco2.py:92 -> Steps.py:201-206 (one-hot at 0); LSA: zeros (Layers.py:356)
    def get_initial_alignment(self, batch_size, key_time, dtype=None):
        cfg = self.engine.cfg
        a = torch.zeros(batch_size, key_time, device="cuda:{}".format(self.engine.device))
        if cfg.attention_type in ("SMA", "BMA"):
            a[:, 0] = 1.0
        return a

    def __call__(self, inputs, training=False, rng: str = "philox", seed: int = 0, step: Optional[int] = None,
                 keep0=None, keep1=None, noise=None, cum_alignment=None):
        return self.call(inputs, training, rng=rng, seed=seed, step=step, keep0=keep0, keep1=keep1, noise=noise,
                         cum_alignment=cum_alignment)

    def call(self, inputs, training=False, rng: str = "philox", seed: int = 0, step: Optional[int] = None,
             keep0=None, keep1=None, noise=None, cum_alignment=None):
        """inputs: [encodings [B,T_v,E], current_mels [B,mel], previous_alignments [B,T_v],
        previous_rnn_states ((h1,c1),(h2,c2))] -> (mel [B,mel*r], stops [B,1], alignments [B,T_v], states).
        For LSA pass/receive the cumulative alignment through ``cum_alignment`` (a mutable list)."""
        if len(inputs) != 4:
            raise ValueError("Unexpected input length")
        encodings, mels, prev_alignment, prev_states = inputs
        eng = self.engine
        cfg = eng.cfg
        if step is None:
            step = self._step
            self._step += 1
        (h1, c1), (h2, c2) = prev_states
        states = torch.stack([torch.as_tensor(s).float().to("cuda:{}".format(eng.device)) for s in (h1, c1, h2, c2)])
        mel_in = torch.as_tensor(mels).float()
        B = mel_in.shape[0]
        k0 = None if keep0 is None else torch.as_tensor(keep0).float().reshape(1, B, -1)
        k1 = None if keep1 is None else torch.as_tensor(keep1).float().reshape(1, B, -1)
        nz = None if noise is None else torch.as_tensor(noise).float().reshape(1, B, -1)
        cum_in = cum_alignment[0] if (cum_alignment is not None and len(cum_alignment)) else None
        want = ("mel", "stop", "alignment", "states") + (("cum_alignment",) if cfg.attention_type == "LSA" else ())
        out = eng.decode(encodings=encodings, teacher_mels=mel_in.reshape(B, 1, cfg.mel_dim), steps=1, rng=rng,
                         keep0=k0, keep1=k1, noise=nz, seed=seed, step_offset=step, init_alignment=prev_alignment,
                         init_cum_alignment=cum_in, init_states=states, want=want, host_outputs=False)
        st = out["states"]
        new_states = ((st[0], st[1]), (st[2], st[3]))
        if cum_alignment is not None and "cum_alignment" in out:
            cum_alignment[:] = [out["cum_alignment"]]
        return out["mel"][:, 0:cfg.step_reduction].reshape(B, -1), out["stop"].reshape(B, 1), \
            out["alignment"][:, 0], new_states


class Decoder:
    """Reference: Modules/Taco2.py:122-232.  The whole tf.while_loop is one persistent-kernel
    launch; ``post_decodings`` (Postnet residual, Taco2.py:230) comes from the
    implicit-GEMM Postnet kernels when the weight pack carries the Postnet variables, and is None
    for a decode-only pack."""

    def __init__(self, engine: Optional[Engine] = None):
        self._engine = engine
        self.layer_Dict = {"Decoder_Step": Decoder_Step(engine)}

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    def __call__(self, inputs, training=False, **kw):
        return self.call(inputs, training, **kw)

    def call(self, inputs, training=False, rng: str = "philox", seed: int = 0, keep0=None, keep1=None, noise=None,
             max_steps: Optional[int] = None, early_stop: bool = False):
        """inputs: [encodings [B,T_v,E], mels [B,T_q,mel]] -> (decodings [B,T*r,mel], post_decodings
        [B,T*r,mel] (None without Postnet variables), stops [B,T], alignments [B,T,T_v])."""
        encodings, mels = inputs
        eng = self.engine
        cfg = eng.cfg
        r = cfg.step_reduction
        if training:
            m = torch.as_tensor(mels) if not isinstance(mels, torch.Tensor) else mels
            teacher = m[:, 0:-1:r, :]                                    # Taco2.py:161
            out = eng.decode(encodings=encodings, teacher_mels=teacher.contiguous(), rng=rng, keep0=keep0,
                             keep1=keep1, noise=noise, seed=seed, steps=max_steps)
        else:
            steps = cfg.max_step // r if max_steps is None else max_steps  # Taco2.py:210-214
            # early_stop (extension): end the loop once every utterance has a negative stop logit - the cut
            # Model.py:380 applies afterwards; the time axis of the outputs is then shorter than Max_Step // r
            out = eng.decode(encodings=encodings, steps=steps, rng=rng, keep0=keep0, keep1=keep1, noise=noise,
                             seed=seed, early_stop=early_stop)
        # The reference's Postnet Sequential inherits training=True from the call context (batch-statistics BatchNorm +
        # Dropout, Taco2.py:144-149): that form is not on the inference path built here, so post_decodings is None
        # under training=True instead of a silently different (inference-form) tensor.
        post = eng.postnet(out["mel"]) if (eng.has_postnet and not training) else None      # Taco2.py:230
        return out["mel"], post, out["stop"], out["alignment"]


class Postnet:
    """Reference: the ``Postnet`` Sequential of Modules/Taco2.py:130-149 together with the residual add of
    :230 (``Postnet(decodings) + decodings``)."""

    def __init__(self, engine: Optional[Engine] = None):
        self._engine = engine

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    def __call__(self, decodings, training=False):
        if training:
            raise NotImplementedError("Postnet: only the inference form (moving-average BatchNorm, no Dropout) is built")
        return self.engine.postnet(decodings)


class Vocoder_Taco1:
    """Reference: Modules/Taco2.py:234-260 - ``Vocoder_Taco1()(inputs= post_mels, training)`` = Dense(Spectrogram_Dim)(CBHG(inputs))
    (CBHG / ConvBank / Highwaynet: Taco2.py:285-434), called on the Postnet output at Model.py:126-129.  [B, T, Mel_Dim] ->
    [B, T, Spectrogram_Dim]; inference form only (moving-average BatchNormalization; the LSTM's recurrent_dropout is 0 in the
    reference's own configuration)."""

    def __init__(self, engine: Optional[Engine] = None):
        self._engine = engine

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    def __call__(self, inputs, training=False):
        return self.call(inputs, training)

    def call(self, inputs, training=False):
        if training:
            raise NotImplementedError("Vocoder_Taco1: only the inference form is built")
        return self.engine.vocoder(inputs)


class Prenet:
    """Reference: Modules/Taco2.py:262-283 - ``Prenet(sizes, dropout_rate)(inputs, training)``: two Dense(relu) + Dropout layers whose
    dropout is ALWAYS on (``training= True   #Always true``).  Inside Decoder_Step the layers are part of the persistent kernel; on
    its own the layer runs through ``gstk_prenet`` with the engine's Decoder_Step/Prenet variables (``sizes`` / ``dropout_rate`` are
    checked against the engine's configuration).  Randomness as for the decoder: ``rng='philox'|'external'|'none'``."""

    def __init__(self, sizes, dropout_rate, engine: Optional[Engine] = None):
        self.sizes = list(sizes)
        self.dropout_rate = dropout_rate
        self._engine = engine

    @property
    def engine(self) -> Engine:
        return self._engine or default_engine()

    def __call__(self, inputs, training=True, **kw):
        return self.call(inputs, training, **kw)

    def call(self, inputs, training=True, rng: str = "philox", seed: int = 0, step: int = 0, keep0=None, keep1=None):
        cfg = self.engine.cfg
        if list(cfg.prenet_sizes) != self.sizes or abs(cfg.prenet_dropout - self.dropout_rate) > 1e-12:
            raise ValueError("Prenet sizes / dropout rate differ from the engine's configuration")
        return self.engine.prenet(inputs, rng=rng, seed=seed, step=step, keep0=keep0, keep1=keep1)
