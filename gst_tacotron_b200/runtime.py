"""Host-side engine over the C ABI: owns one GstkHandle (device + packed weights + workspace) and
moves tensors across the boundary without copies (device tensors are borrowed by pointer – torch
tensors directly, anything else that speaks DLPack, e.g. TF2 eager tensors, via
``torch.from_dlpack``; numpy arrays are passed as host pointers and staged by the library).

PyTorch is used for device memory and streams only; every FLOP of the hot path runs in
libgsttaco.so."""
from __future__ import annotations

import ctypes as C
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Dict, List, Mapping, Optional

import numpy as np
import torch

from . import _lib
from .hparams import HotPathConfig
from .weights import check_weights, encoder_spec, vocoder_spec, postnet_spec, weight_spec


def _to_tensor(x):
    """torch.Tensor / numpy / DLPack-capable object -> (float32 contiguous torch.Tensor or ndarray)."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.float32)
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    else:
        return np.ascontiguousarray(np.asarray(x), dtype=np.float32)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    return x.ctypes.data


class Engine:
    """One handle = one device.  Not thread-safe (the reference's model call is single-threaded,
    Model.py:249-255)."""

    def __init__(self, cfg: HotPathConfig, weights: Optional[Mapping[str, np.ndarray]] = None, device: int = 0):
        cfg.validate()
        self.cfg = cfg
        self.device = int(device)
        self._lib = _lib.load()
        c = _lib.GstkConfig()
        c.version = _lib.GSTK_VERSION
        c.device = self.device
        c.mel_dim = cfg.mel_dim
        c.step_reduction = cfg.step_reduction
        c.prenet0, c.prenet1 = cfg.prenet_sizes
        c.attention_size = cfg.attention_size
        c.attention_type = _lib.ATT[cfg.attention_type]
        c.lstm0, c.lstm1 = cfg.lstm_sizes
        c.enc_dim = cfg.enc_dim
        c.gst_use = int(cfg.gst_use)
        c.ref_layers = len(cfg.ref_filters)
        for i, (f, k, s) in enumerate(zip(cfg.ref_filters, cfg.ref_kernel, cfg.ref_strides)):
            c.ref_filters[i], c.ref_kernel[i], c.ref_stride[i] = f, k, s
        c.ref_gru = cfg.ref_gru_size
        c.ref_dense = cfg.ref_dense_size
        c.n_tokens = cfg.n_tokens
        c.token_dim = cfg.token_dim
        c.style_heads = cfg.style_heads
        c.style_size = cfg.style_size
        c.lsa_filters = cfg.lsa_filters
        c.lsa_kernel = cfg.lsa_kernel
        c.lsa_cumulate = int(cfg.lsa_cumulate)
        c.lsa_smoothing = int(cfg.lsa_smoothing)
        c.precision = _lib.PREC[cfg.precision]
        c.prenet_dropout = cfg.prenet_dropout
        c.sigmoid_noise = cfg.sigmoid_noise
        h = C.c_void_p()
        rc = self._lib.gstk_create(C.byref(c), C.byref(h))
        _lib.raise_for(rc, None)
        self._h = h
        self.has_postnet = False
        self.has_encoder = False
        self.has_vocoder = False
        if weights is not None:
            self.load_weights(weights)

    # ------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gstk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _lib.raise_for(rc, self._h)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def load_weights(self, weights: Mapping[str, np.ndarray]) -> None:
        """Checkpoint restore (Model.py:267-276) for the hot-path variables."""
        check_weights(self.cfg, weights)
        names = list(weight_spec(self.cfg).keys())
        # the Postnet variables (Taco2.py:130-147) are optional: a pack without them decodes, and postnet() then fails
        # with GSTK_ENOWEIGHTS
        # the Postnet / text Encoder variables are optional in the same way (encoder() needs the latter)
        for attr, spec in (("has_postnet", postnet_spec(self.cfg)), ("has_encoder", encoder_spec(self.cfg)),
                           ("has_vocoder", vocoder_spec(self.cfg))):
            if any(n in weights for n in spec):
                for n, shape in spec.items():
                    if n not in weights:
                        raise KeyError("missing variable {}".format(n))
                    if tuple(weights[n].shape) != tuple(shape):
                        raise ValueError("variable {} has shape {}, expected {}".format(n, weights[n].shape, shape))
                names += list(spec.keys())
                setattr(self, attr, True)
        descs = (_lib.GstkTensorDesc * len(names))()
        keep = []
        for i, n in enumerate(names):
            a = np.ascontiguousarray(weights[n], dtype=np.float32)
            keep.append(a)
            descs[i].name = n.encode()
            descs[i].data = a.ctypes.data
            descs[i].ndim = a.ndim
            for k, s in enumerate(a.shape):
                descs[i].shape[k] = s
        self._check(self._lib.gstk_load_weights(self._h, descs, len(names)))

    # ------------------------------------------------------------------------------------
    def _alloc(self, shape, host: bool):
        if host:
            return np.empty(shape, dtype=np.float32)
        return torch.empty(shape, dtype=torch.float32, device="cuda:{}".format(self.device))

    def decode(self, encodings=None, enc_text=None, gst=None, teacher_mels=None, steps: Optional[int] = None,
               rng: str = "none", keep0=None, keep1=None, noise=None, seed: int = 0, step_offset: int = 0,
               row_offset: int = 0, init_mel=None, init_alignment=None, init_cum_alignment=None, init_states=None,
               want=("mel", "stop", "alignment"), host_outputs: Optional[bool] = None, early_stop: bool = False,
               out_buffers: Optional[Dict[str, torch.Tensor]] = None, kernel: str = "auto") -> Dict[str, object]:
        """Run `steps` decoder steps (Decoder.call's loop, Taco2.py:182-226, without the Postnet).

        teacher_mels given  => training=True semantics: step t consumes teacher_mels[:, t]
                               (already sliced ``mels[:, 0:-1:r]``, Taco2.py:161).
        teacher_mels None   => free running from `init_mel` (zeros by default).
        Returns a dict with the requested outputs among mel [B,T*r,mel], stop [B,T],
        alignment [B,T,Tv], states [4,B,U], cum_alignment [B,Tv], context [B,A].

        early_stop=True (free running): the loop ends once every utterance has produced a negative stop logit - the cut the
        reference's caller applies afterwards (Model.py:380).  The time axis of mel / stop / alignment is then truncated to
        ``steps_done`` and the dict also carries ``stop_index`` [B] (first step with stop < 0, T if none) and ``steps_done``.
        ``want`` may name "stop_index" on its own to get the indices of a full-length decode.

        kernel (bf16 engines): "auto" - the small-batch latency kernel for free-running SMA decodes of batch <= 16 (key_time <= 256,
        no early_stop), the batch-256 kernel otherwise; "batch" / "small" / "dataflow" pin one (GstkDecodeArgs::kernel).  The
        kernels agree within the bf16 tolerance, not bit for bit: pin "batch" when pieces of one job must equal the whole."""
        cfg = self.cfg
        enc_t = _to_tensor(encodings)
        text_t, gst_t = _to_tensor(enc_text), _to_tensor(gst)
        src = enc_t if enc_t is not None else text_t
        if src is None:
            raise ValueError("pass encodings, or enc_text + gst")
        if src.ndim != 3:
            raise ValueError("encodings / enc_text must be [batch, key_time, channels]")
        B, Tv = int(src.shape[0]), int(src.shape[1])
        # the C ABI takes plain pointers: the channel counts it will read are checked here
        if enc_t is not None and int(enc_t.shape[2]) != cfg.enc_dim:
            raise ValueError("encodings must have {} channels (GST channels first), got {}".format(cfg.enc_dim, int(enc_t.shape[2])))
        if enc_t is None:
            if gst_t is None:
                raise ValueError("enc_text needs gst")
            if int(text_t.shape[2]) != cfg.text_dim or tuple(gst_t.shape) != (B, cfg.style_size):
                raise ValueError("enc_text must be [B, T_v, {}] and gst [B, {}]".format(cfg.text_dim, cfg.style_size))
        teach = _to_tensor(teacher_mels)
        if teach is not None:
            T = int(teach.shape[1]) if steps is None else min(int(steps), int(teach.shape[1]))
        else:
            T = cfg.max_step // cfg.step_reduction if steps is None else int(steps)
        if host_outputs is None:
            host_outputs = not isinstance(src, torch.Tensor) or not src.is_cuda
        a = _lib.GstkDecodeArgs()
        a.batch, a.key_time, a.steps = B, Tv, T
        a.mode = _lib.MODE_TEACHER if teach is not None else _lib.MODE_FREE
        a.rng_mode = _lib.RNG[rng]
        a.seed = seed
        a.step_offset = step_offset
        a.row_offset = row_offset
        holders = [enc_t, text_t, gst_t, teach]
        a.encodings, a.enc_text, a.gst = _ptr(enc_t), _ptr(text_t), _ptr(gst_t)
        if teach is not None:
            a.teacher_mels = _ptr(teach)
            a.teacher_stride_b = int(teach.shape[1]) * cfg.mel_dim
            a.teacher_stride_t = cfg.mel_dim
        for name, val in (("keep0", keep0), ("keep1", keep1), ("noise", noise), ("init_mel", init_mel),
                          ("init_alignment", init_alignment), ("init_cum_alignment", init_cum_alignment),
                          ("init_states", init_states)):
            t = _to_tensor(val)
            holders.append(t)
            setattr(a, name, _ptr(t))
        out: Dict[str, object] = {}
        shapes = {
            "mel": (B, T * cfg.step_reduction, cfg.mel_dim), "stop": (B, T), "alignment": (B, T, Tv),
            "states": (4, B, cfg.lstm_sizes[0]), "cum_alignment": (B, Tv), "context": (B, cfg.attention_size),
        }
        fields = {"mel": "out_mel", "stop": "out_stop", "alignment": "out_alignment", "states": "out_states",
                  "cum_alignment": "out_cum_alignment", "context": "out_context"}
        want_idx = early_stop or "stop_index" in want
        for k in want:
            if k == "stop_index":
                continue
            buf = None if out_buffers is None else out_buffers.get(k)
            if buf is None:
                buf = self._alloc(shapes[k], host_outputs)
            elif tuple(buf.shape) != tuple(shapes[k]) or buf.dtype != torch.float32 or not buf.is_contiguous():
                raise ValueError("out_buffers[{!r}] must be a contiguous float32 tensor of shape {}".format(k, shapes[k]))
            out[k] = buf
            setattr(a, fields[k], _ptr(buf))
        if want_idx:
            if teach is not None and early_stop:
                raise ValueError("early_stop applies to free-running decodes only")
            idx = np.zeros(B, np.int32)
            done = np.zeros(1, np.int32)
            a.early_stop = 1 if early_stop else 0
            a.out_stop_index = idx.ctypes.data
            a.out_steps_done = done.ctypes.data
        a.kernel = _lib.KERNEL[kernel]
        a.stream = self._stream()
        self._check(self._lib.gstk_decode(self._h, C.byref(a)))
        del holders
        if want_idx:
            out["stop_index"] = idx
            out["steps_done"] = int(done[0])
            if early_stop:
                n = int(done[0])
                for k, per in (("mel", cfg.step_reduction), ("stop", 1), ("alignment", 1)):
                    if k in out:
                        out[k] = out[k][:, :n * per]
        return out

    def gst(self, mels, lengths, drop_first: bool = True, want=("gst",), host_outputs: Optional[bool] = None):
        """Style_Token_Layer.call (drop_first=True, GST.py:91-109) / Reference_Encoder.call."""
        cfg = self.cfg
        m = _to_tensor(mels)
        B, frames = int(m.shape[0]), int(m.shape[1])
        if isinstance(lengths, torch.Tensor):
            ln = lengths.to(torch.int32).contiguous()
        else:
            ln = np.ascontiguousarray(np.asarray(lengths), dtype=np.int32)
        if host_outputs is None:
            host_outputs = not isinstance(m, torch.Tensor) or not m.is_cuda
        a = _lib.GstkGstArgs()
        a.batch, a.frames, a.drop_first = B, frames, int(drop_first)
        a.mels, a.lengths = _ptr(m), _ptr(ln)
        shapes = {"gst": (B, cfg.style_size), "ref": (B, cfg.ref_dense_size), "attention": (B, cfg.n_tokens)}
        fields = {"gst": "out_gst", "ref": "out_ref", "attention": "out_attention"}
        out = {}
        for k in want:
            buf = self._alloc(shapes[k], host_outputs)
            out[k] = buf
            setattr(a, fields[k], _ptr(buf))
        a.stream = self._stream()
        self._check(self._lib.gstk_gst(self._h, C.byref(a)))
        return out

    def postnet(self, decodings, host_outputs: Optional[bool] = None):
        """post_decodings = Postnet(decodings) + decodings (Taco2.py:230); decodings [B, T*r, mel]."""
        cfg = self.cfg
        d = _to_tensor(decodings)
        B, T = int(d.shape[0]), int(d.shape[1])
        if int(d.shape[2]) != cfg.mel_dim:
            raise ValueError("decodings must have Mel_Dim channels")
        layers = cfg.postnet_layers
        if any(s != 1 for (_f, _k, s, _t) in layers):
            raise ValueError("Postnet strides other than 1 are not supported (the reference's residual add needs stride 1)")
        if host_outputs is None:
            host_outputs = not isinstance(d, torch.Tensor) or not d.is_cuda
        a = _lib.GstkPostnetArgs()
        a.batch, a.frames, a.n_layers = B, T, len(layers)
        for i, (f, k, _s, th) in enumerate(layers):
            a.filters[i], a.kernel[i], a.use_tanh[i] = f, k, int(th)
        out = self._alloc((B, T, cfg.mel_dim), host_outputs)
        a.decodings, a.out_post = _ptr(d), _ptr(out)
        a.stream = self._stream()
        self._check(self._lib.gstk_postnet(self._h, C.byref(a)))
        return out

    def encoder(self, tokens, host_outputs: Optional[bool] = None):
        """Encoder.call (Taco2.py:47-51): tokens [B, T_v] integers -> [B, T_v, 2 * Encoder.RNN.Size]."""
        cfg = self.cfg
        if any(s != 1 for s in cfg.encoder_strides):
            raise ValueError("Encoder conv strides other than 1 are not supported")
        if isinstance(tokens, torch.Tensor):
            tk = tokens.to(torch.int32).contiguous()
        elif hasattr(tokens, "__dlpack__") and not isinstance(tokens, np.ndarray):
            tk = torch.from_dlpack(tokens).to(torch.int32).contiguous()
        else:
            tk = np.ascontiguousarray(np.asarray(tokens), dtype=np.int32)
        if tk.ndim != 2:
            raise ValueError("tokens must be [batch, key_time]")
        B, Tv = int(tk.shape[0]), int(tk.shape[1])
        if host_outputs is None:
            host_outputs = not isinstance(tk, torch.Tensor) or not tk.is_cuda
        a = _lib.GstkEncoderArgs()
        a.batch, a.key_time, a.vocab, a.embedding = B, Tv, cfg.vocab_size, cfg.encoder_embedding
        a.n_layers, a.rnn_size = len(cfg.encoder_filters), cfg.encoder_rnn_size
        for i, (f, k) in enumerate(zip(cfg.encoder_filters, cfg.encoder_kernel)):
            a.filters[i], a.kernel[i] = f, k
        out = self._alloc((B, Tv, 2 * cfg.encoder_rnn_size), host_outputs)
        a.tokens, a.out = _ptr(tk), _ptr(out)
        a.stream = self._stream()
        self._check(self._lib.gstk_encoder(self._h, C.byref(a)))
        return out

    def prenet(self, inputs, rng: str = "philox", seed: int = 0, step: int = 0, row_offset: int = 0, keep0=None, keep1=None,
               host_outputs: Optional[bool] = None):
        """Prenet.call (Taco2.py:282-283) on its own: inputs [..., Mel_Dim] -> [..., Prenet.Size[-1]], dropout always on
        (rng="none" switches it off, "external" takes keep0 / keep1 masks of {0, 1}, "philox" draws the decoder's streams)."""
        cfg = self.cfg
        x = _to_tensor(inputs)
        if int(x.shape[-1]) != cfg.mel_dim:
            raise ValueError("inputs must have Mel_Dim channels")
        lead = tuple(int(v) for v in x.shape[:-1])
        rows = int(np.prod(lead)) if lead else 1
        if host_outputs is None:
            host_outputs = not isinstance(x, torch.Tensor) or not x.is_cuda
        a = _lib.GstkPrenetArgs()
        a.rows, a.rng_mode, a.seed, a.step, a.row_offset = rows, _lib.RNG[rng], seed, step, row_offset
        k0, k1 = _to_tensor(keep0), _to_tensor(keep1)
        out = self._alloc(lead + (cfg.prenet_sizes[-1],), host_outputs)
        a.inputs, a.keep0, a.keep1, a.out = _ptr(x), _ptr(k0), _ptr(k1), _ptr(out)
        a.stream = self._stream()
        self._check(self._lib.gstk_prenet(self._h, C.byref(a)))
        return out

    def vocoder(self, mels, host_outputs: Optional[bool] = None):
        """Vocoder_Taco1.call (Taco2.py:258-260): mels [B, T, Mel_Dim] (the Postnet output, Model.py:126-129) ->
        linear spectrogram [B, T, Spectrogram_Dim]."""
        cfg = self.cfg
        m = _to_tensor(mels)
        if m.ndim != 3 or int(m.shape[2]) != cfg.mel_dim:
            raise ValueError("mels must be [batch, frames, Mel_Dim]")
        B, T = int(m.shape[0]), int(m.shape[1])
        if host_outputs is None:
            host_outputs = not isinstance(m, torch.Tensor) or not m.is_cuda
        a = _lib.GstkVocoderArgs()
        a.batch, a.frames = B, T
        a.bank_count, a.bank_filters = cfg.voc_bank_count, cfg.voc_bank_filters
        a.pool_size, a.pool_strides = cfg.voc_pool_size, cfg.voc_pool_strides
        a.n_proj = len(cfg.voc_proj_filters)
        if a.n_proj > 8 or len(cfg.voc_proj_kernel) != a.n_proj:
            raise ValueError("Vocoder: 1..8 projection conv layers with one kernel size each")
        for i, (f, k) in enumerate(zip(cfg.voc_proj_filters, cfg.voc_proj_kernel)):
            a.proj_filters[i], a.proj_kernel[i] = f, k
        a.highway_count, a.highway_size = cfg.voc_highway_count, cfg.voc_highway_size
        a.rnn_size, a.spectrogram_dim = cfg.voc_rnn_size, cfg.spectrogram_dim
        out = self._alloc((B, T, cfg.spectrogram_dim), host_outputs)
        a.mels, a.out = _ptr(m), _ptr(out)
        a.stream = self._stream()
        self._check(self._lib.gstk_vocoder(self._h, C.byref(a)))
        return out

    def griffin_lim(self, spectrogram, lengths=None, iters: Optional[int] = None, rng: str = "philox", seed: int = 0,
                    init_uniform=None, row_offset: int = 0, ref_level_db: float = 20.0, power: float = 1.5,
                    max_abs_value: Optional[float] = None, preemphasis: float = 0.97, hop_length: Optional[int] = None,
                    win_length: Optional[int] = None, host_outputs: Optional[bool] = None):
        """Audio.inv_spectrogram (Audio.py:23-27) for a batch: spectrogram [B, T, num_freq] as the vocoder returns it
        (the reference transposes one utterance at a time, Model.py:413) -> waveforms [B, hop * (T - 1)]; utterance b holds
        hop * (lengths[b] - 1) samples, zeros after.  rng="external": ``init_uniform`` [B, T, num_freq] stands for
        np.random.rand (Audio.py:61)."""
        cfg = self.cfg
        sp = _to_tensor(spectrogram)
        if sp.ndim != 3:
            raise ValueError("spectrogram must be [batch, frames, num_freq]")
        B, T, F = int(sp.shape[0]), int(sp.shape[1]), int(sp.shape[2])
        if host_outputs is None:
            host_outputs = not isinstance(sp, torch.Tensor) or not sp.is_cuda
        a = _lib.GstkGriffinLimArgs()
        a.batch, a.frames, a.num_freq = B, T, F
        a.hop_length = cfg.frame_shift if hop_length is None else int(hop_length)
        a.win_length = cfg.frame_length if win_length is None else int(win_length)
        a.iters = cfg.griffin_lim_iters if iters is None else int(iters)
        a.rng_mode = _lib.RNG[rng]
        a.row_offset, a.seed = row_offset, seed
        a.ref_level_db, a.power, a.preemphasis = ref_level_db, power, preemphasis
        a.max_abs_value = -1.0 if max_abs_value is None else float(max_abs_value)
        ln = None
        if lengths is not None:
            ln = lengths.to(torch.int32).contiguous() if isinstance(lengths, torch.Tensor) else \
                np.ascontiguousarray(np.asarray(lengths), dtype=np.int32)
        un = _to_tensor(init_uniform)
        out = self._alloc((B, a.hop_length * (T - 1)), host_outputs)
        a.spectrogram, a.lengths, a.init_uniform, a.out_wav = _ptr(sp), _ptr(ln), _ptr(un), _ptr(out)
        a.stream = self._stream()
        self._check(self._lib.gstk_griffin_lim(self._h, C.byref(a)))
        return out

    def inference(self, tokens, mels_for_gst, mel_lengths_for_gst, steps: Optional[int] = None, rng: str = "philox",
                  seed: int = 0, keep0=None, keep1=None, noise=None, host_outputs: Optional[bool] = None, wav: bool = False):
        """The reference's ``Inference`` functional model (Model.py:108-129, called at :249-253): Encoder(tokens) ->
        Style_Token_Layer([mels_for_gst, lengths]) -> GST_Concated_Encoder (folded into the value projection) -> Decoder
        free-running for Max_Step // Step_Reduction steps -> Postnet residual -> Vocoder_Taco1 (when its variables are
        loaded).  Everything stays on the device between the stages.  Returns dict(mel, post_mel, stop, alignment,
        spectrogram, encodings, gst).  wav=True adds what Export_Inference does next (Model.py:380,412-420): every
        utterance cut at its first negative stop logit and turned into a waveform by Griffin-Lim - ``stop_index`` [B] and
        ``wav`` [B, Frame_Shift * (T - 1)] (zeros beyond Frame_Shift * (max(1, stop_index) * r - 1) samples)."""
        dev = "cuda:{}".format(self.device)
        tk = tokens if isinstance(tokens, torch.Tensor) else torch.as_tensor(np.asarray(tokens))
        if host_outputs is None:
            host_outputs = not tk.is_cuda
        enc = self.encoder(tk.to(dev), host_outputs=False)
        m = _to_tensor(mels_for_gst)
        m = m.to(dev) if isinstance(m, torch.Tensor) else torch.as_tensor(m, device=dev)
        ln = torch.as_tensor(np.asarray(mel_lengths_for_gst) if not isinstance(mel_lengths_for_gst, torch.Tensor)
                             else mel_lengths_for_gst).to(dev)
        g = self.gst(m, ln, want=("gst",), host_outputs=False)["gst"]
        out = self.decode(enc_text=enc, gst=g, steps=steps, rng=rng, seed=seed, keep0=keep0, keep1=keep1, noise=noise,
                          host_outputs=False)
        post = self.postnet(out["mel"], host_outputs=False) if self.has_postnet else None
        spec = self.vocoder(post, host_outputs=False) if (self.has_vocoder and post is not None) else None
        res = {"mel": out["mel"], "post_mel": post, "stop": out["stop"], "alignment": out["alignment"], "spectrogram": spec,
               "encodings": enc, "gst": g}
        if wav:
            if spec is None:
                raise ValueError("wav=True needs the Postnet and Vocoder_Taco1 variables")
            stop = out["stop"]
            neg = stop < 0
            idx = torch.where(neg.any(1), neg.to(torch.int32).argmax(1), torch.zeros_like(neg[:, 0], dtype=torch.int64))  # np.argmax(stop < 0)
            ln = (torch.clamp(idx, min=1) * self.cfg.step_reduction).to(torch.int32)
            res["stop_index"] = idx.to(torch.int32)
            res["wav"] = self.griffin_lim(spec, lengths=ln, rng="philox", seed=seed, max_abs_value=self.cfg.max_abs_mel,
                                          host_outputs=False)
        if host_outputs:
            res = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}
        return res

    def mha(self, query, value, q_kernel, q_bias, v_kernel, v_bias, ln_gamma, ln_beta, heads: int,
            host_outputs: Optional[bool] = None):
        q, v = _to_tensor(query), _to_tensor(value)
        ws = [_to_tensor(w) for w in (q_kernel, q_bias, v_kernel, v_bias, ln_gamma, ln_beta)]
        B, tq, dq = [int(s) for s in q.shape]
        tv, dv = int(v.shape[1]), int(v.shape[2])
        size = int(ws[0].shape[1])
        if host_outputs is None:
            host_outputs = not isinstance(q, torch.Tensor) or not q.is_cuda
        a = _lib.GstkMhaArgs()
        a.batch, a.tq, a.tv, a.dq, a.dv, a.size, a.heads = B, tq, tv, dq, dv, size, heads
        a.query, a.value = _ptr(q), _ptr(v)
        (a.q_kernel, a.q_bias, a.v_kernel, a.v_bias, a.ln_gamma, a.ln_beta) = [_ptr(w) for w in ws]
        out = self._alloc((B, tq, size), host_outputs)
        att = self._alloc((B, tq, tv), host_outputs)
        a.out, a.out_attention = _ptr(out), _ptr(att)
        a.stream = self._stream()
        self._check(self._lib.gstk_mha(self._h, C.byref(a)))
        return out, att

    def concat_encoder(self, enc_text, gst, host_outputs: Optional[bool] = None):
        e, g = _to_tensor(enc_text), _to_tensor(gst)
        B, Tv = int(e.shape[0]), int(e.shape[1])
        if host_outputs is None:
            host_outputs = not isinstance(e, torch.Tensor) or not e.is_cuda
        out = self._alloc((B, Tv, self.cfg.enc_dim), host_outputs)
        self._check(self._lib.gstk_concat_encoder(self._h, _ptr(e), _ptr(g), _ptr(out), B, Tv, self._stream()))
        return out

    def synchronize(self):
        self._check(self._lib.gstk_synchronize(self._h, self._stream()))

    @property
    def launch_count(self) -> int:
        return int(self._lib.gstk_launch_count(self._h))

    def phase_profile(self):
        """[n_ctas, 16] clock64 ticks of the last bf16 decode launch per phase (diagnostics)."""
        buf = np.zeros((256, 16), dtype=np.uint64)
        n = C.c_int32(0)
        self._check(self._lib.gstk_get_phase_profile(self._h, buf.ctypes.data, 256, C.byref(n)))
        return buf[: n.value]

    def last_kernel_ms(self) -> float:
        return float(self._lib.gstk_last_kernel_ms(self._h))


class AutoEngine:
    """``precision="auto"``: the tensor-core engine wherever it applies, the exact fp32 engine otherwise - stated, never silent.

    * configurations the tensor-core decoder does not cover (LSTM sizes other than 1024 / 1024, prenet + attention width other than
      384: rejected by gstk_create) get an fp32 engine from the start; ``precision_used`` says which, and a ``UserWarning`` names the
      reason the library gave;
    * a call the tensor-core engine rejects for its SHAPE (key_time beyond its shared-memory budget) is repeated on an fp32 engine
      that is created on first need from the same weights (``fallback_calls`` counts them).
    Everything else is the wrapped engine's (attribute access is delegated)."""

    def __init__(self, cfg: HotPathConfig, weights: Optional[Mapping[str, np.ndarray]] = None, device: int = 0):
        import copy
        import warnings
        self._weights, self._device = weights, device
        self._cfg32 = copy.copy(cfg)
        self._cfg32.precision = "fp32"
        self._fallback: Optional[Engine] = None
        self.fallback_calls = 0
        cfg16 = copy.copy(cfg)
        cfg16.precision = "bf16"
        try:
            self._main = Engine(cfg16, weights, device=device)
            self.precision_used = "bf16"
        except ValueError as e:
            if "bf16 tensor-core path" not in str(e):
                raise
            warnings.warn("gst_tacotron_b200: tensor-core mode does not cover this configuration ({}); using the fp32 engine".format(e))
            self._main = Engine(self._cfg32, weights, device=device)
            self.precision_used = "fp32"

    def _fp32(self) -> "Engine":
        if self.precision_used == "fp32":
            return self._main
        if self._fallback is None:
            self._fallback = Engine(self._cfg32, self._weights, device=self._device)
        return self._fallback

    def decode(self, *args, **kwargs):
        try:
            return self._main.decode(*args, **kwargs)
        except ValueError as e:
            if self.precision_used == "fp32" or "key_time" not in str(e):
                raise
            self.fallback_calls += 1
            kwargs.pop("kernel", None)
            return self._fp32().decode(*args, **kwargs)

    def close(self):
        self._main.close()
        if self._fallback is not None:
            self._fallback.close()
            self._fallback = None

    def __getattr__(self, name):
        return getattr(self._main, name)


class EnginePool:
    """``depth`` engines on ONE device, each with its own weights image, activation slots, CUDA stream and worker thread.
    Requests submitted back to back alternate between them: while one engine's persistent decoder kernel owns the SMs, the other
    engine's host->device input copies (and its previous request's device->host output copies) run on the copy engines, so a
    stream of requests with host buffers costs the kernels' time, not kernels + PCIe.  The reference has no equivalent (its
    Inference_Step is one synchronous Keras call, Model.py:249-255); this is the serving loop around the drop-in.

        pool = EnginePool(cfg, weights, device=0, depth=2)
        futs = [pool.submit(lambda eng, x=x: eng.decode(enc_text=x.text, gst=x.gst, steps=T, rng="philox")) for x in requests]
        outs = [f.result() for f in futs]

    ``fn`` runs on the worker thread of the engine it is given, with that engine's stream current; results come back in
    submission order through the futures.  Every engine is an ordinary :class:`Engine` (one request at a time each)."""

    def __init__(self, cfg: HotPathConfig, weights: Optional[Mapping[str, np.ndarray]] = None, device: int = 0, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.device = device
        self.engines: List[Engine] = [Engine(cfg, weights, device=device) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(depth)]
        self._workers = [ThreadPoolExecutor(max_workers=1, thread_name_prefix="gstk-engine-{}".format(i)) for i in range(depth)]
        self._next = 0
        self._lock = threading.Lock()

    def submit(self, fn: Callable[[Engine], object], engine: Optional[int] = None) -> Future:
        with self._lock:
            k = self._next if engine is None else engine
            if engine is None:
                self._next = (self._next + 1) % len(self.engines)

        def run():
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(self.streams[k]):
                return fn(self.engines[k])

        return self._workers[k].submit(run)

    def synchronize(self):
        for k in range(len(self.engines)):
            self.submit(lambda e: e.synchronize(), engine=k).result()

    def close(self):
        for w in self._workers:
            w.shutdown(wait=True)
        for e in self.engines:
            e.close()
        self.engines = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
