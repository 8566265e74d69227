"""Weight pack for the decode hot path.

Variable names and shapes mirror the Keras variables the reference creates
(reference: Modules/Taco2.py:59-94,262-283, Modules/Attention/Steps.py:65-86,
Modules/GST.py:12-45,72-89, Modules/Attention/Layers.py:162-168,259-277).  The pack is a plain
``{path: float32 ndarray}`` dict; paths use the reference's ``layer_Dict`` keys so a dict pulled
out of a ``tf.train.Checkpoint`` (Model.py:186-189) can be matched by suffix + shape with
:func:`from_named_arrays`.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping

import numpy as np

from .hparams import HotPathConfig

DEC = "Decoder/Decoder_Step"
GST = "Style_Token_Layer"
REF = GST + "/Reference_Encoder"
POST = "Decoder/Postnet"
ENC = "Encoder"
VOC = "Vocoder_Taco1"


def weight_spec(cfg: HotPathConfig) -> "OrderedDict[str, tuple]":
    """Ordered {path: shape} for every variable on the hot path."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    p0, p1 = cfg.prenet_sizes
    A = cfg.attention_size
    u0, u1 = cfg.lstm_sizes
    # Prenet (Taco2.py:270-279): Sequential of Dense(relu)+Dropout => dense, dense_1
    s[DEC + "/Prenet/dense/kernel"] = (cfg.mel_dim, p0)
    s[DEC + "/Prenet/dense/bias"] = (p0,)
    s[DEC + "/Prenet/dense_1/kernel"] = (p0, p1)
    s[DEC + "/Prenet/dense_1/bias"] = (p1,)
    # Attention (Steps.py:65-86). 'Key' Dense is constructed but never built on the 3-input path.
    s[DEC + "/Attention/Query/kernel"] = (p1, A)
    s[DEC + "/Attention/Query/bias"] = (A,)
    s[DEC + "/Attention/Value/kernel"] = (cfg.enc_dim, A)
    s[DEC + "/Attention/Value/bias"] = (A,)
    if cfg.attention_type in ("SMA", "BMA"):
        s[DEC + "/Attention/attention_v"] = (A,)
        s[DEC + "/Attention/attention_score_bias"] = ()
    else:  # LSA (Layers.py:310-341)
        s[DEC + "/Attention/Alignment_Conv/kernel"] = (cfg.lsa_kernel, 1, cfg.lsa_filters)
        s[DEC + "/Attention/Alignment_Conv/bias"] = (cfg.lsa_filters,)
        s[DEC + "/Attention/Alignment_Dense/kernel"] = (cfg.lsa_filters, A)
        s[DEC + "/Attention/Alignment_Dense/bias"] = (A,)
        s[DEC + "/Attention/bias"] = (A,)
    # StackedRNNCells[LSTMCell, LSTMCell] (Taco2.py:77-85)
    s[DEC + "/RNN/cell_0/kernel"] = (p1 + A, 4 * u0)
    s[DEC + "/RNN/cell_0/recurrent_kernel"] = (u0, 4 * u0)
    s[DEC + "/RNN/cell_0/bias"] = (4 * u0,)
    s[DEC + "/RNN/cell_1/kernel"] = (u0, 4 * u1)
    s[DEC + "/RNN/cell_1/recurrent_kernel"] = (u1, 4 * u1)
    s[DEC + "/RNN/cell_1/bias"] = (4 * u1,)
    # Projection (Taco2.py:87-89)
    s[DEC + "/Projection/kernel"] = (u1 + A, cfg.proj_dim)
    s[DEC + "/Projection/bias"] = (cfg.proj_dim,)
    if cfg.gst_use:
        cin = 1
        for i, cout in enumerate(cfg.ref_filters):
            k = cfg.ref_kernel[i]
            base = REF + "/Conv2D_{}".format(i)
            s[base + "/conv2d/kernel"] = (k, k, cin, cout)  # HWIO
            s[base + "/batch_normalization/gamma"] = (cout,)
            s[base + "/batch_normalization/beta"] = (cout,)
            s[base + "/batch_normalization/moving_mean"] = (cout,)
            s[base + "/batch_normalization/moving_variance"] = (cout,)
            cin = cout
        mel_w = cfg.mel_dim
        for st in cfg.ref_strides:
            mel_w = -(-mel_w // st)
        gin = mel_w * cfg.ref_filters[-1]
        g = cfg.ref_gru_size
        s[REF + "/RNN/kernel"] = (gin, 3 * g)
        s[REF + "/RNN/recurrent_kernel"] = (g, 3 * g)
        s[REF + "/RNN/bias"] = (2, 3 * g)  # reset_after=True layout
        s[REF + "/Dense/kernel"] = (g, cfg.ref_dense_size)
        s[REF + "/Dense/bias"] = (cfg.ref_dense_size,)
        S = cfg.style_size
        s[GST + "/Attention/Query/kernel"] = (cfg.ref_dense_size, S)
        s[GST + "/Attention/Query/bias"] = (S,)
        s[GST + "/Attention/Value/kernel"] = (cfg.token_dim, S)
        s[GST + "/Attention/Value/bias"] = (S,)
        s[GST + "/Attention/Layer_Normalization/beta"] = (S,)
        s[GST + "/Attention/Layer_Normalization/gamma"] = (S,)
        s[GST + "/gst_tokens"] = (cfg.n_tokens, cfg.token_dim)
    return s


def postnet_spec(cfg: HotPathConfig) -> "OrderedDict[str, tuple]":
    """Ordered {path: shape} of the Postnet variables (Taco2.py:130-147): per layer a bias-free Conv1D kernel
    [kernel_size, in, out] and BatchNormalization statistics.  Kept apart from :func:`weight_spec`: the
    decode loop does not need them, and a pack without them stays valid."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    cin = cfg.mel_dim
    for i, (cout, k, _stride, _tanh) in enumerate(cfg.postnet_layers):
        s[POST + "/conv1d_{}/kernel".format(i)] = (k, cin, cout)
        for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[POST + "/batch_normalization_{}/{}".format(i, leaf)] = (cout,)
        cin = cout
    return s


def init_postnet_weights(cfg: HotPathConfig, seed: int = 4321) -> Dict[str, np.ndarray]:
    """Random Postnet variables (own generator: the decode pack of init_weights(seed) is unchanged)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in postnet_spec(cfg).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf == "kernel":
            w = _glorot(rng, shape)
        elif leaf in ("gamma", "moving_variance"):
            w = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        else:
            w = rng.normal(0.0, 0.1, size=shape).astype(np.float32)
        out[name] = np.ascontiguousarray(w, dtype=np.float32).reshape(shape)
    return out


def encoder_spec(cfg: HotPathConfig) -> "OrderedDict[str, tuple]":
    """Ordered {path: shape} of the text Encoder variables (Taco2.py:16-45): Embedding table, per conv layer a bias-free
    Conv1D kernel [k, in, out] + BatchNormalization statistics, and the two LSTM cells of the Bidirectional wrapper
    (kernel [in, 4u], recurrent_kernel [u, 4u], bias [4u]; gate blocks i|f|c|o)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s[ENC + "/embedding/embeddings"] = (cfg.vocab_size, cfg.encoder_embedding)
    cin = cfg.encoder_embedding
    for i, (cout, k) in enumerate(zip(cfg.encoder_filters, cfg.encoder_kernel)):
        s[ENC + "/conv1d_{}/kernel".format(i)] = (k, cin, cout)
        for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[ENC + "/batch_normalization_{}/{}".format(i, leaf)] = (cout,)
        cin = cout
    u = cfg.encoder_rnn_size
    for d in ("forward_lstm", "backward_lstm"):
        s[ENC + "/bidirectional/{}/lstm_cell/kernel".format(d)] = (cin, 4 * u)
        s[ENC + "/bidirectional/{}/lstm_cell/recurrent_kernel".format(d)] = (u, 4 * u)
        s[ENC + "/bidirectional/{}/lstm_cell/bias".format(d)] = (4 * u,)
    return s


def init_encoder_weights(cfg: HotPathConfig, seed: int = 2468, bias_scale: float = 0.05) -> Dict[str, np.ndarray]:
    """Random Encoder variables (own generator).  Embedding ~ U(-0.05, 0.05) (Keras 'uniform'), forget-gate bias + 1."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in encoder_spec(cfg).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf == "embeddings":
            w = rng.uniform(-0.05, 0.05, size=shape).astype(np.float32) * 20.0   # spread so that the conv stack is exercised
        elif leaf in ("kernel", "recurrent_kernel"):
            w = _glorot(rng, shape)
        elif leaf in ("gamma", "moving_variance"):
            w = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "bias":
            w = rng.normal(0.0, bias_scale, size=shape).astype(np.float32)
            u = shape[0] // 4
            w[u:2 * u] += 1.0
        else:
            w = rng.normal(0.0, 0.1, size=shape).astype(np.float32)
        out[name] = np.ascontiguousarray(w, dtype=np.float32).reshape(shape)
    return out


def vocoder_spec(cfg: HotPathConfig) -> "OrderedDict[str, tuple]":
    """Ordered {path: shape} of the Vocoder_Taco1 variables (Taco2.py:234-260; CBHG :285-385, ConvBank :388-414, Highwaynet
    :416-434) in construction order:
      ConvBank_i        bias-free Conv1D(bank_filters, kernel_size i + 1) + BatchNormalization         (:396-405)
      Conv1D_Projection bias-free Conv1D + BatchNormalization per entry of Conv1D.Filters, then a Dense(Mel_Dim) when the last
                        filter count differs from the input's channel count                                (:328-345)
      Highwaynet        Dense(size) when the input's channel count differs from it, then `count` Highwaynet layers of a
                        Dense_Relu and a Dense_Sigmoid                                                      (:347-356, 419-429)
      RNN               Bidirectional(LSTM(rnn_size)): kernel [in, 4u], recurrent_kernel [u, 4u], bias [4u]   (:358-362)
      Dense             Dense(Spectrogram_Dim)                                                               (:253-255)"""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    mel = cfg.mel_dim
    for i in range(cfg.voc_bank_count):
        s[VOC + "/CBHG/ConvBank_{}/conv1d/kernel".format(i)] = (i + 1, mel, cfg.voc_bank_filters)
        for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[VOC + "/CBHG/ConvBank_{}/batch_normalization/{}".format(i, leaf)] = (cfg.voc_bank_filters,)
    cin = cfg.voc_bank_count * cfg.voc_bank_filters
    for i, (cout, k) in enumerate(zip(cfg.voc_proj_filters, cfg.voc_proj_kernel)):
        s[VOC + "/CBHG/Conv1D_Projection/conv1d_{}/kernel".format(i)] = (k, cin, cout)
        for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[VOC + "/CBHG/Conv1D_Projection/batch_normalization_{}/{}".format(i, leaf)] = (cout,)
        cin = cout
    if cin != mel:
        s[VOC + "/CBHG/Conv1D_Projection/dense/kernel"] = (cin, mel)
        s[VOC + "/CBHG/Conv1D_Projection/dense/bias"] = (mel,)
    hs = cfg.voc_highway_size
    if mel != hs:
        s[VOC + "/CBHG/Highwaynet/dense/kernel"] = (mel, hs)
        s[VOC + "/CBHG/Highwaynet/dense/bias"] = (hs,)
    for i in range(cfg.voc_highway_count):
        for nm in ("Dense_Relu", "Dense_Sigmoid"):
            s[VOC + "/CBHG/Highwaynet/highwaynet_{}/{}/kernel".format(i, nm)] = (hs, hs)
            s[VOC + "/CBHG/Highwaynet/highwaynet_{}/{}/bias".format(i, nm)] = (hs,)
    u = cfg.voc_rnn_size
    for d in ("forward_lstm", "backward_lstm"):
        s[VOC + "/CBHG/RNN/{}/lstm_cell/kernel".format(d)] = (hs, 4 * u)
        s[VOC + "/CBHG/RNN/{}/lstm_cell/recurrent_kernel".format(d)] = (u, 4 * u)
        s[VOC + "/CBHG/RNN/{}/lstm_cell/bias".format(d)] = (4 * u,)
    s[VOC + "/Dense/kernel"] = (2 * u, cfg.spectrogram_dim)
    s[VOC + "/Dense/bias"] = (cfg.spectrogram_dim,)
    return s


def init_vocoder_weights(cfg: HotPathConfig, seed: int = 1357, bias_scale: float = 0.05) -> Dict[str, np.ndarray]:
    """Random Vocoder_Taco1 variables (own generator; forget-gate bias + 1 as Keras' unit_forget_bias)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in vocoder_spec(cfg).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf in ("kernel", "recurrent_kernel"):
            w = _glorot(rng, shape)
        elif leaf in ("gamma", "moving_variance"):
            w = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "bias":
            w = rng.normal(0.0, bias_scale, size=shape).astype(np.float32)
            if "lstm_cell" in name:
                u = shape[0] // 4
                w[u:2 * u] += 1.0
        else:
            w = rng.normal(0.0, 0.1, size=shape).astype(np.float32)
        out[name] = np.ascontiguousarray(w, dtype=np.float32).reshape(shape)
    return out


def postnet_pack(cfg: HotPathConfig, conv_kernels, bn_params) -> Dict[str, np.ndarray]:
    """Postnet variables from the reference's layer objects IN LAYER ORDER (Keras auto-names such as ``conv1d_7`` depend on
    how many layers were built before, so they are matched by position, not by name):
    ``conv_kernels[i]`` = i-th Conv1D kernel [k, in, out] of ``Decoder.layer_Dict['Postnet'].layers`` (Taco2.py:131-149),
    ``bn_params[i]`` = (gamma, beta, moving_mean, moving_variance) of the i-th BatchNormalization."""
    spec = postnet_spec(cfg)
    n = len(cfg.postnet_layers)
    if len(conv_kernels) != n or len(bn_params) != n:
        raise ValueError("expected {} Postnet conv / batch-norm layers".format(n))
    out: Dict[str, np.ndarray] = {}
    for i in range(n):
        out[POST + "/conv1d_{}/kernel".format(i)] = conv_kernels[i]
        for leaf, v in zip(("gamma", "beta", "moving_mean", "moving_variance"), bn_params[i]):
            out[POST + "/batch_normalization_{}/{}".format(i, leaf)] = v
    return _checked(spec, out)


def encoder_pack(cfg: HotPathConfig, embeddings, conv_kernels, bn_params, forward_cell, backward_cell) -> Dict[str, np.ndarray]:
    """Encoder variables from the reference's layer objects in layer order (``Encoder.layer.layers``, Taco2.py:16-45):
    Embedding table, Conv1D kernels, BatchNormalization (gamma, beta, moving_mean, moving_variance) tuples, and
    (kernel, recurrent_kernel, bias) of ``Bidirectional.forward_layer.cell`` / ``.backward_layer.cell``."""
    spec = encoder_spec(cfg)
    n = len(cfg.encoder_filters)
    if len(conv_kernels) != n or len(bn_params) != n:
        raise ValueError("expected {} Encoder conv / batch-norm layers".format(n))
    out: Dict[str, np.ndarray] = {ENC + "/embedding/embeddings": embeddings}
    for i in range(n):
        out[ENC + "/conv1d_{}/kernel".format(i)] = conv_kernels[i]
        for leaf, v in zip(("gamma", "beta", "moving_mean", "moving_variance"), bn_params[i]):
            out[ENC + "/batch_normalization_{}/{}".format(i, leaf)] = v
    for d, cell in (("forward_lstm", forward_cell), ("backward_lstm", backward_cell)):
        for leaf, v in zip(("kernel", "recurrent_kernel", "bias"), cell):
            out[ENC + "/bidirectional/{}/lstm_cell/{}".format(d, leaf)] = v
    return _checked(spec, out)


def vocoder_pack(cfg: HotPathConfig, bank, projection, projection_dense, highway_dense, highways, forward_cell, backward_cell,
                 dense) -> Dict[str, np.ndarray]:
    """Vocoder_Taco1 variables from the reference's layer objects in construction order (Keras auto-names are not stable):
      bank[i]           = (Conv1D kernel [i + 1, mel, filters], (gamma, beta, moving_mean, moving_variance)) of
                          ``CBHG.layer_Dict['ConvBank'].layer_Dict['ConvBank_i'].layers`` (Taco2.py:396-405)
      projection[i]     = the same pair for the i-th Conv1D / BatchNormalization of ``layer_Dict['Conv1D_Projection']`` (:328-340)
      projection_dense  = (kernel, bias) of its trailing Dense(Mel_Dim), or None when the last filter count equals Mel_Dim (:341-345)
      highway_dense     = (kernel, bias) of the Dense(size) in front of the Highwaynet layers, or None (:347-351)
      highways[i]       = ((kernel, bias) of Dense_Relu, (kernel, bias) of Dense_Sigmoid) of the i-th Highwaynet (:419-429)
      forward_cell / backward_cell = (kernel, recurrent_kernel, bias) of ``layer_Dict['RNN'].forward_layer.cell`` / ``.backward_layer.cell``
      dense             = (kernel, bias) of ``Vocoder_Taco1.layer_Dict['Dense']`` (:253-255)"""
    spec = vocoder_spec(cfg)
    if len(bank) != cfg.voc_bank_count or len(projection) != len(cfg.voc_proj_filters) or len(highways) != cfg.voc_highway_count:
        raise ValueError("expected {} conv-bank, {} projection and {} Highwaynet layers".format(
            cfg.voc_bank_count, len(cfg.voc_proj_filters), cfg.voc_highway_count))
    out: Dict[str, np.ndarray] = {}
    bn_leaves = ("gamma", "beta", "moving_mean", "moving_variance")
    for i, (kernel, bn) in enumerate(bank):
        out[VOC + "/CBHG/ConvBank_{}/conv1d/kernel".format(i)] = kernel
        for leaf, v in zip(bn_leaves, bn):
            out[VOC + "/CBHG/ConvBank_{}/batch_normalization/{}".format(i, leaf)] = v
    for i, (kernel, bn) in enumerate(projection):
        out[VOC + "/CBHG/Conv1D_Projection/conv1d_{}/kernel".format(i)] = kernel
        for leaf, v in zip(bn_leaves, bn):
            out[VOC + "/CBHG/Conv1D_Projection/batch_normalization_{}/{}".format(i, leaf)] = v
    for key, pair in ((VOC + "/CBHG/Conv1D_Projection/dense/", projection_dense), (VOC + "/CBHG/Highwaynet/dense/", highway_dense),
                      (VOC + "/Dense/", dense)):
        if (key + "kernel" in spec) != (pair is not None):
            raise ValueError("{} {} for this configuration".format(key, "is required" if pair is None else "does not exist"))
        if pair is not None:
            out[key + "kernel"], out[key + "bias"] = pair
    for i, (relu, sig) in enumerate(highways):
        for nm, pair in (("Dense_Relu", relu), ("Dense_Sigmoid", sig)):
            out[VOC + "/CBHG/Highwaynet/highwaynet_{}/{}/kernel".format(i, nm)], out[VOC + "/CBHG/Highwaynet/highwaynet_{}/{}/bias".format(i, nm)] = pair
    for d, cell in (("forward_lstm", forward_cell), ("backward_lstm", backward_cell)):
        for leaf, v in zip(("kernel", "recurrent_kernel", "bias"), cell):
            out[VOC + "/CBHG/RNN/{}/lstm_cell/{}".format(d, leaf)] = v
    return _checked(spec, out)


def _checked(spec, named) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for name, shape in spec.items():
        a = np.ascontiguousarray(np.asarray(named[name]), dtype=np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError("variable {} has shape {}, expected {}".format(name, a.shape, shape))
        out[name] = a
    return out


def _glorot(rng: np.random.Generator, shape) -> np.ndarray:
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:  # conv: receptive field * channels
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_weights(cfg: HotPathConfig, seed: int = 1234, bias_scale: float = 0.0) -> Dict[str, np.ndarray]:
    """Keras-like random initialisation: glorot-uniform kernels, zero biases
    (``bias_scale`` > 0 draws N(0, bias_scale) biases instead so bias handling is exercised), LSTM
    forget-gate bias 1, non-trivial BatchNorm statistics, truncated-normal style tokens."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in weight_spec(cfg).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf in ("kernel", "recurrent_kernel", "attention_v"):
            w = _glorot(rng, shape)
        elif leaf == "gamma" and "batch_normalization" in name:
            w = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "moving_variance":
            w = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf in ("beta", "moving_mean") and "batch_normalization" in name:
            w = rng.normal(0.0, 0.1, size=shape).astype(np.float32)
        elif leaf == "gamma":
            w = np.ones(shape, np.float32)
        elif leaf == "beta":
            w = np.zeros(shape, np.float32)
        elif leaf == "gst_tokens":
            w = np.clip(rng.normal(0.0, 0.5, size=shape), -1.0, 1.0).astype(np.float32)
        elif leaf in ("bias", "attention_score_bias"):
            if bias_scale > 0.0:
                w = rng.normal(0.0, bias_scale, size=shape).astype(np.float32)
            else:
                w = np.zeros(shape, np.float32)
            if "/RNN/cell_" in name:  # unit_forget_bias=True
                u = shape[0] // 4
                w = w.copy()
                w[u:2 * u] += 1.0
        else:
            raise KeyError(name)
        out[name] = np.ascontiguousarray(w, dtype=np.float32).reshape(shape)
    return out


def check_weights(cfg: HotPathConfig, weights: Mapping[str, np.ndarray]) -> None:
    for name, shape in weight_spec(cfg).items():
        if name not in weights:
            raise KeyError("missing variable {}".format(name))
        if tuple(weights[name].shape) != tuple(shape):
            raise ValueError("variable {} has shape {}, expected {}".format(name, weights[name].shape, shape))


def from_named_arrays(cfg: HotPathConfig, named: Mapping[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Build a pack from a ``{path: ndarray}`` dict whose paths need only *end with* the
    canonical names (case-insensitive, ':0' suffix ignored) – e.g. a dump of the reference's
    checkpoint object graph, whose exact key strings could not be verified without TensorFlow."""
    spec = weight_spec(cfg)
    norm = {}
    for k, v in named.items():
        kk = k.replace(":0", "").replace(".", "/").lower()
        norm[kk] = np.asarray(v)
    out = {}
    for name, shape in spec.items():
        key = name.lower()
        hits = [k for k in norm if k == key or k.endswith("/" + key)]
        hits = [k for k in hits if tuple(norm[k].shape) == tuple(shape)]
        if len(hits) != 1:
            raise KeyError("variable {}: {} candidates with shape {}".format(name, len(hits), shape))
        out[name] = np.ascontiguousarray(norm[hits[0]], dtype=np.float32).reshape(shape)
    return out


def save_npz(path: str, weights: Mapping[str, np.ndarray]) -> None:
    np.savez(path, **{k.replace("/", "|"): v for k, v in weights.items()})


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k.replace("|", "/"): np.ascontiguousarray(z[k], dtype=np.float32).reshape(z[k].shape) for k in z.files}
