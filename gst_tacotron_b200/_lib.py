"""ctypes binding of libgsttaco.so (include/gstk.h).  There is no fallback: if the shared
library has not been built, or no B200 is present, the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSTK_LIB_PATH", os.path.join(HERE, "libgsttaco.so"))  # override: A/B builds during development

GSTK_VERSION = 1
(GSTK_OK, GSTK_EINVAL, GSTK_ENODEVICE, GSTK_ECUDA, GSTK_ENOWEIGHTS, GSTK_ENOTIMPL, GSTK_ETIMEOUT) = range(7)
ATT = {"SMA": 0, "BMA": 1, "LSA": 2}
PREC = {"fp32": 0, "bf16": 1}
RNG = {"none": 0, "external": 1, "philox": 2}
MODE_FREE, MODE_TEACHER = 0, 1
KERNEL = {"auto": 0, "batch": 1, "small": 2, "dataflow": 3}   # GstkDecodeArgs::kernel

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)


class GstkConfig(C.Structure):
    _fields_ = [
        ("version", C.c_int32), ("device", C.c_int32), ("mel_dim", C.c_int32), ("step_reduction", C.c_int32),
        ("prenet0", C.c_int32), ("prenet1", C.c_int32), ("attention_size", C.c_int32), ("attention_type", C.c_int32),
        ("lstm0", C.c_int32), ("lstm1", C.c_int32), ("enc_dim", C.c_int32), ("gst_use", C.c_int32),
        ("ref_layers", C.c_int32), ("ref_filters", C.c_int32 * 8), ("ref_kernel", C.c_int32 * 8),
        ("ref_stride", C.c_int32 * 8), ("ref_gru", C.c_int32), ("ref_dense", C.c_int32), ("n_tokens", C.c_int32),
        ("token_dim", C.c_int32), ("style_heads", C.c_int32), ("style_size", C.c_int32),
        ("lsa_filters", C.c_int32), ("lsa_kernel", C.c_int32), ("lsa_cumulate", C.c_int32), ("lsa_smoothing", C.c_int32),
        ("precision", C.c_int32), ("prenet_dropout", C.c_float), ("sigmoid_noise", C.c_float),
        ("reserved", C.c_int32 * 8),
    ]


class GstkTensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class GstkDecodeArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("key_time", C.c_int32), ("steps", C.c_int32), ("mode", C.c_int32),
        ("rng_mode", C.c_int32), ("pad0", C.c_int32), ("seed", C.c_uint64), ("step_offset", C.c_uint32),
        ("row_offset", C.c_uint32),
        ("encodings", C.c_void_p), ("enc_text", C.c_void_p), ("gst", C.c_void_p), ("teacher_mels", C.c_void_p),
        ("teacher_stride_b", C.c_int64), ("teacher_stride_t", C.c_int64),
        ("keep0", C.c_void_p), ("keep1", C.c_void_p), ("noise", C.c_void_p),
        ("init_mel", C.c_void_p), ("init_alignment", C.c_void_p), ("init_cum_alignment", C.c_void_p),
        ("init_states", C.c_void_p),
        ("out_mel", C.c_void_p), ("out_stop", C.c_void_p), ("out_alignment", C.c_void_p), ("out_states", C.c_void_p),
        ("out_cum_alignment", C.c_void_p), ("out_context", C.c_void_p),
        ("stream", C.c_void_p),
        ("early_stop", C.c_int32), ("kernel", C.c_int32), ("out_stop_index", C.c_void_p), ("out_steps_done", C.c_void_p),
        ("reserved", C.c_int32 * 2),
    ]


class GstkGstArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("frames", C.c_int32), ("drop_first", C.c_int32), ("pad0", C.c_int32),
        ("mels", C.c_void_p), ("lengths", C.c_void_p), ("out_gst", C.c_void_p), ("out_ref", C.c_void_p),
        ("out_attention", C.c_void_p), ("stream", C.c_void_p), ("reserved", C.c_int32 * 8),
    ]


class GstkPostnetArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("frames", C.c_int32), ("n_layers", C.c_int32), ("pad0", C.c_int32),
        ("filters", C.c_int32 * 8), ("kernel", C.c_int32 * 8), ("use_tanh", C.c_int32 * 8),
        ("decodings", C.c_void_p), ("out_post", C.c_void_p), ("stream", C.c_void_p), ("reserved", C.c_int32 * 8),
    ]


class GstkEncoderArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("key_time", C.c_int32), ("vocab", C.c_int32), ("embedding", C.c_int32),
        ("n_layers", C.c_int32), ("rnn_size", C.c_int32), ("filters", C.c_int32 * 8), ("kernel", C.c_int32 * 8),
        ("tokens", C.c_void_p), ("out", C.c_void_p), ("stream", C.c_void_p), ("reserved", C.c_int32 * 8),
    ]


class GstkVocoderArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("frames", C.c_int32), ("bank_count", C.c_int32), ("bank_filters", C.c_int32),
        ("pool_size", C.c_int32), ("pool_strides", C.c_int32), ("n_proj", C.c_int32), ("highway_count", C.c_int32),
        ("highway_size", C.c_int32), ("rnn_size", C.c_int32), ("spectrogram_dim", C.c_int32), ("pad0", C.c_int32),
        ("proj_filters", C.c_int32 * 8), ("proj_kernel", C.c_int32 * 8),
        ("mels", C.c_void_p), ("out", C.c_void_p), ("stream", C.c_void_p), ("reserved", C.c_int32 * 8),
    ]


class GstkGriffinLimArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("frames", C.c_int32), ("num_freq", C.c_int32), ("hop_length", C.c_int32),
        ("win_length", C.c_int32), ("iters", C.c_int32), ("rng_mode", C.c_int32), ("row_offset", C.c_int32),
        ("ref_level_db", C.c_float), ("power", C.c_float), ("max_abs_value", C.c_float), ("preemphasis", C.c_float),
        ("seed", C.c_uint64),
        ("spectrogram", C.c_void_p), ("lengths", C.c_void_p), ("init_uniform", C.c_void_p), ("out_wav", C.c_void_p),
        ("stream", C.c_void_p), ("reserved", C.c_int32 * 8),
    ]


class GstkPrenetArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("step", C.c_uint32), ("row_offset", C.c_int32),
        ("inputs", C.c_void_p), ("keep0", C.c_void_p), ("keep1", C.c_void_p), ("out", C.c_void_p), ("stream", C.c_void_p),
        ("reserved", C.c_int32 * 4),
    ]


class GstkMhaArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("tq", C.c_int32), ("tv", C.c_int32), ("dq", C.c_int32), ("dv", C.c_int32),
        ("size", C.c_int32), ("heads", C.c_int32), ("pad0", C.c_int32),
        ("query", C.c_void_p), ("value", C.c_void_p), ("q_kernel", C.c_void_p), ("q_bias", C.c_void_p),
        ("v_kernel", C.c_void_p), ("v_bias", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("out", C.c_void_p), ("out_attention", C.c_void_p), ("stream", C.c_void_p),
    ]


class GstkAttentionArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("key_time", C.c_int32), ("query_dim", C.c_int32), ("value_dim", C.c_int32),
        ("key_dim", C.c_int32), ("size", C.c_int32), ("type", C.c_int32), ("pad0", C.c_int32),
        ("sigmoid_noise", C.c_float), ("pad1", C.c_float),
        ("query", C.c_void_p), ("value", C.c_void_p), ("key", C.c_void_p), ("prev_alignment", C.c_void_p),
        ("noise", C.c_void_p), ("q_kernel", C.c_void_p), ("q_bias", C.c_void_p), ("v_kernel", C.c_void_p),
        ("v_bias", C.c_void_p), ("k_kernel", C.c_void_p), ("k_bias", C.c_void_p), ("attention_v", C.c_void_p),
        ("attention_score_bias", C.c_void_p), ("out_context", C.c_void_p), ("out_alignment", C.c_void_p),
        ("stream", C.c_void_p),
    ]


EXPORTS = {
    "gstk_version": (C.c_int, []),
    "gstk_create": (C.c_int, [C.POINTER(GstkConfig), C.POINTER(C.c_void_p)]),
    "gstk_destroy": (C.c_int, [C.c_void_p]),
    "gstk_load_weights": (C.c_int, [C.c_void_p, C.POINTER(GstkTensorDesc), C.c_int32]),
    "gstk_decode": (C.c_int, [C.c_void_p, C.POINTER(GstkDecodeArgs)]),
    "gstk_gst": (C.c_int, [C.c_void_p, C.POINTER(GstkGstArgs)]),
    "gstk_postnet": (C.c_int, [C.c_void_p, C.POINTER(GstkPostnetArgs)]),
    "gstk_encoder": (C.c_int, [C.c_void_p, C.POINTER(GstkEncoderArgs)]),
    "gstk_vocoder": (C.c_int, [C.c_void_p, C.POINTER(GstkVocoderArgs)]),
    "gstk_griffin_lim": (C.c_int, [C.c_void_p, C.POINTER(GstkGriffinLimArgs)]),
    "gstk_prenet": (C.c_int, [C.c_void_p, C.POINTER(GstkPrenetArgs)]),
    "gstk_mha": (C.c_int, [C.c_void_p, C.POINTER(GstkMhaArgs)]),
    "gstk_attention_step": (C.c_int, [C.c_void_p, C.POINTER(GstkAttentionArgs)]),
    "gstk_concat_encoder": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "gstk_synchronize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gstk_launch_count": (C.c_int64, [C.c_void_p]),
    "gstk_last_kernel_ms": (C.c_float, [C.c_void_p]),
    "gstk_last_error": (C.c_char_p, [C.c_void_p]),
    "gstk_get_phase_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "gstk_selftest_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib = None


class GstkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libgsttaco error {}: {}".format(code, msg))
        self.code = code
        self.msg = msg


def load() -> C.CDLL:
    """Load libgsttaco.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libgsttaco.so is not built ({}). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python gst_tacotron_b200/build.py`. gst_tacotron_b200 has no CPU fallback.".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.gstk_version() != GSTK_VERSION:
        raise RuntimeError("libgsttaco.so version mismatch")
    _lib = lib
    return lib


def raise_for(code: int, handle) -> None:
    """Map C status codes onto the exception types the reference raises (SURVEY.md 8b 'Errors')."""
    if code == GSTK_OK:
        return
    msg = load().gstk_last_error(handle)
    msg = msg.decode() if msg else ""
    if code == GSTK_EINVAL:
        raise ValueError(msg)
    if code == GSTK_ENOTIMPL:
        raise NotImplementedError(msg)
    raise GstkError(code, msg)
