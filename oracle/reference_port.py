"""CPU ORACLE (test infrastructure, NOT product code).

A CPU restatement, in torch-CPU tensors (float64 by default), of the reference's decode hot
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this file.  The product path (``gst_tacotron_b200``) never
does, and fails loudly when its CUDA library is missing.

PARITY PINNING: the reference holds no golden vectors or tests for this path (SURVEY.md
section 4) and TensorFlow is not installable here, so the Keras *primitive* semantics (Dense,
LSTMCell, GRU reset_after, Conv2D 'same' padding, BatchNormalization inference form) are
restated from TF 2.x behaviour as listed in SURVEY.md section 8c.  Everything that lives in the
reference's own source (decoder step wiring, attention scoring, monotonic probability
functions, safe_cumprod, style-token multi-head attention, Layer_Norm, GST concat, decoder
loop conventions) is additionally pinned by ``tests/golden/*.npz``, which were produced by
executing the reference's unmodified Python sources from /root/reference on top of a minimal
torch-backed TensorFlow API shim (``oracle/tf_shim`` + ``oracle/make_golden.py``).

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DEC = "Decoder/Decoder_Step"
GST = "Style_Token_Layer"
REF = GST + "/Reference_Encoder"
POST = "Decoder/Postnet"
ENC = "Encoder"
VOC = "Vocoder_Taco1"

F32_TINY = float(np.finfo(np.float32).tiny)  # Steps.py:197 (np.finfo(dtype).tiny for float32)


def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x)).to(dtype)


def to_torch(weights: Mapping[str, np.ndarray], dtype=torch.float64) -> Dict[str, torch.Tensor]:
    return {k: _t(v, dtype) for k, v in weights.items()}


# ----------------------------------------------------------------------------------------
# Keras primitives (semantics: SURVEY.md section 8c)
# ----------------------------------------------------------------------------------------
def dense(x, kernel, bias):
    """tf.keras.layers.Dense: y = x . kernel + bias, kernel [in, out]."""
    return x @ kernel + bias


def dropout_apply(x, keep, rate: float):
    """tf.keras.layers.Dropout in training mode: kept units are scaled by 1/(1-rate)."""
    if rate <= 0.0:
        return x
    return x * keep * (1.0 / (1.0 - rate))


def lstm_cell(x, h, c, kernel, recurrent_kernel, bias):
    """tf.keras.layers.LSTMCell (TF2 defaults): z = x.W + h.U + b, gate order i,f,c,o,
    recurrent_activation sigmoid, activation tanh.  Returns (h', c')."""
    z = x @ kernel + h @ recurrent_kernel + bias
    i, f, g, o = torch.chunk(z, 4, dim=-1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


def gru_sequence(x, kernel, recurrent_kernel, bias):
    """tf.keras.layers.GRU(return_sequences=True), TF2 default reset_after=True, gate order z,r,h,
    bias [2, 3u] (input bias, recurrent bias), zero initial state, no mask."""
    B, T, _ = x.shape
    u = recurrent_kernel.shape[0]
    h = torch.zeros(B, u, dtype=x.dtype)
    xs = x @ kernel + bias[0]
    outs = []
    for t in range(T):
        hh = h @ recurrent_kernel + bias[1]
        xz, xr, xh = torch.chunk(xs[:, t], 3, dim=-1)
        hz, hr, hn = torch.chunk(hh, 3, dim=-1)
        z = torch.sigmoid(xz + hz)
        r = torch.sigmoid(xr + hr)
        cand = torch.tanh(xh + r * hn)
        h = z * h + (1.0 - z) * cand
        outs.append(h)
    return torch.stack(outs, dim=1)


def conv2d_same_nhwc(x, kernel_hwio, stride: int):
    """tf.keras.layers.Conv2D(padding='same', use_bias=False) on NHWC input with TF's asymmetric
    'same' padding: out = ceil(in/stride); pad_total = max((out-1)*stride + k - in, 0);
    pad_before = pad_total // 2."""
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    H, W = x.shape[1], x.shape[2]

    def pads(n, k):
        out = -(-n // stride)
        tot = max((out - 1) * stride + k - n, 0)
        return tot // 2, tot - tot // 2

    pt, pb = pads(H, kh)
    pl, pr = pads(W, kw)
    xn = x.permute(0, 3, 1, 2)  # NCHW
    xn = F.pad(xn, (pl, pr, pt, pb))
    w = kernel_hwio.permute(3, 2, 0, 1)  # OIHW
    y = F.conv2d(xn, w, stride=stride)
    return y.permute(0, 2, 3, 1)


def batchnorm_inference(x, gamma, beta, mean, var, eps: float = 1e-3):
    """tf.keras.layers.BatchNormalization, inference form, axis=-1, epsilon=1e-3 (Keras default)."""
    return gamma * (x - mean) / torch.sqrt(var + eps) + beta


# ----------------------------------------------------------------------------------------
# Decoder step (Modules/Taco2.py:53-120, 262-283; Modules/Attention/Steps.py:51-229)
# ----------------------------------------------------------------------------------------
def prenet(W, cfg, x, keep0, keep1):
    """Prenet.call, Taco2.py:282-283: Sequential[Dense(relu), Dropout]x2, training=True always."""
    h = torch.relu(dense(x, W[DEC + "/Prenet/dense/kernel"], W[DEC + "/Prenet/dense/bias"]))
    h = dropout_apply(h, keep0, cfg.prenet_dropout)
    h = torch.relu(dense(h, W[DEC + "/Prenet/dense_1/kernel"], W[DEC + "/Prenet/dense_1/bias"]))
    h = dropout_apply(h, keep1, cfg.prenet_dropout)
    return h


def value_projection(W, encodings):
    """Steps.py:123 `value = Dense_V(value)`; key = value on the 3-input path (Steps.py:124).
    Loop-invariant: the reference recomputes it every step, the result is the same tensor."""
    return dense(encodings, W[DEC + "/Attention/Value/kernel"], W[DEC + "/Attention/Value/bias"])


def monotonic_scores(W, query, key):
    """_calculate_scores, Steps.py:138-152 (normalize=False):
    e_j = sum_a v_a * tanh(q_a + k_ja) + score_bias.  query [B,A], key [B,Tv,A] -> [B,Tv]."""
    v = W[DEC + "/Attention/attention_v"]
    b = W[DEC + "/Attention/attention_score_bias"]
    return torch.sum(v * torch.tanh(query[:, None, :] + key), dim=-1) + b


def sma_probability(score, prev_alignment, noise, sigmoid_noise: float):
    """StepwiseMonotonicAttention._monotonic_probability_fn, Steps.py:215-229."""
    if sigmoid_noise > 0.0:
        score = score + sigmoid_noise * noise
    p = torch.sigmoid(score)
    moved = prev_alignment[:, :-1] * (1.0 - p[:, :-1])
    pad = torch.zeros_like(p[:, :1])
    return prev_alignment * p + torch.cat([pad, moved], dim=-1)


def safe_cumprod_exclusive(x):
    """safe_cumprod(x, axis, exclusive=True), Steps.py:183-198:
    exp(cumsum_exclusive(log(clip(x, tiny, 1))))."""
    lg = torch.log(torch.clamp(x, F32_TINY, 1.0))
    cs = torch.cumsum(lg, dim=-1) - lg
    return torch.exp(cs)


def bma_probability(score, prev_alignment, noise, sigmoid_noise: float):
    """BahdanauMonotonicAttention._monotonic_probability_fn, Steps.py:168-180."""
    if sigmoid_noise > 0.0:
        score = score + sigmoid_noise * noise
    p = torch.sigmoid(score)
    cp = safe_cumprod_exclusive(1.0 - p)
    return p * cp * torch.cumsum(prev_alignment / torch.clamp(cp, 1e-10, 1.0), dim=-1)


def conv1d_same_nwc(x, kernel_wio, bias, stride: int = 1):
    """tf.keras.layers.Conv1D(padding='same') with bias on [B, W, C] input (TF 'same' padding)."""
    k = kernel_wio.shape[0]
    n = x.shape[1]
    out = -(-n // stride)
    tot = max((out - 1) * stride + k - n, 0)
    xn = F.pad(x.permute(0, 2, 1), (tot // 2, tot - tot // 2))
    y = F.conv1d(xn, kernel_wio.permute(2, 1, 0), bias=bias, stride=stride)
    return y.permute(0, 2, 1)


def lsa_alignment(W, cfg, query, key, location_source):
    """Step form of LocationSensitiveAttention (Layers.py:359-369, 393-424):
    loc = Dense_A(Conv1D(prev[..., None])); e_j = sum_a tanh(q + k_j + loc_j + bias)_a (no v vector,
    scale 1); alignment = softmax(e) or smoothing-normalised sigmoid."""
    loc = conv1d_same_nwc(location_source[:, :, None], W[DEC + "/Attention/Alignment_Conv/kernel"],
                          W[DEC + "/Attention/Alignment_Conv/bias"])
    loc = dense(loc, W[DEC + "/Attention/Alignment_Dense/kernel"], W[DEC + "/Attention/Alignment_Dense/bias"])
    e = torch.sum(torch.tanh(query[:, None, :] + key + loc + W[DEC + "/Attention/bias"]), dim=-1)
    if cfg.lsa_smoothing:
        s = torch.sigmoid(e)
        return s / torch.sum(s, dim=-1, keepdim=True)
    return torch.softmax(e, dim=-1)


def initial_alignment(cfg, batch: int, key_time: int, dtype):
    """Steps.py:201-206: one-hot at index 0 (SMA/BMA); LSA starts from zeros (Layers.py:356)."""
    a = torch.zeros(batch, key_time, dtype=dtype)
    if cfg.attention_type in ("SMA", "BMA"):
        a[:, 0] = 1.0
    return a


def decoder_step(W, cfg, values, mel_in, prev_alignment, states, keep0, keep1, noise,
                 lsa_cum=None):
    """Decoder_Step.call, Taco2.py:96-120, with the value projection passed in (``values`` =
    V' [B,Tv,A]).  states = ((h1,c1),(h2,c2)).  Returns (mel [B,80r], stop [B,1], alignment,
    states, context)."""
    p = prenet(W, cfg, mel_in, keep0, keep1)                                       # Taco2.py:106
    q = dense(p, W[DEC + "/Attention/Query/kernel"], W[DEC + "/Attention/Query/bias"])  # Steps.py:122
    if cfg.attention_type == "SMA":
        e = monotonic_scores(W, q, values)
        align = sma_probability(e, prev_alignment, noise, cfg.sigmoid_noise)
    elif cfg.attention_type == "BMA":
        e = monotonic_scores(W, q, values)
        align = bma_probability(e, prev_alignment, noise, cfg.sigmoid_noise)
    else:
        src = lsa_cum if cfg.lsa_cumulate else prev_alignment
        align = lsa_alignment(W, cfg, q, values, src)
    ctx = torch.einsum("bt,bta->ba", align, values)                                # Steps.py:164
    (h1, c1), (h2, c2) = states
    x1 = torch.cat([p, ctx], dim=-1)                                               # Taco2.py:110
    h1, c1 = lstm_cell(x1, h1, c1, W[DEC + "/RNN/cell_0/kernel"], W[DEC + "/RNN/cell_0/recurrent_kernel"],
                       W[DEC + "/RNN/cell_0/bias"])
    h2, c2 = lstm_cell(h1, h2, c2, W[DEC + "/RNN/cell_1/kernel"], W[DEC + "/RNN/cell_1/recurrent_kernel"],
                       W[DEC + "/RNN/cell_1/bias"])
    y = dense(torch.cat([h2, ctx], dim=-1), W[DEC + "/Projection/kernel"], W[DEC + "/Projection/bias"])
    mel, stop = y[:, :-1], y[:, -1:]                                               # Taco2.py:114-118
    return mel, stop, align, ((h1, c1), (h2, c2)), ctx


def decoder_loop(weights, cfg, encodings, mels=None, training: bool = False, steps: Optional[int] = None,
                 keep0=None, keep1=None, noise=None, dtype=torch.float64,
                 reproject_every_step: bool = False):
    """Decoder.call without the Postnet, Taco2.py:153-228.

    encodings [B,Tv,E]; ``mels`` [B,Tq,mel] (teacher, includes the initial zero frame; sliced
    ``mels[:, 0:-1:r]`` as Taco2.py:161).  ``training`` selects teacher forcing.  Explicit
    randomness: keep0/keep1 [T,B,p] in {0,1}, noise [T,B,Tv] ~ N(0,1) (None => all-keep / zeros).
    Returns dict(decodings [B,T*r,mel], stops [B,T], alignments [B,T,Tv], states).
    ``reproject_every_step`` recomputes the value projection inside the loop exactly as the
    reference does (Steps.py:123) – same numbers, reference-like cost (used by the CPU baseline).
    """
    W = weights if all(isinstance(v, torch.Tensor) and v.dtype == dtype for v in weights.values()) \
        else to_torch(weights, dtype)
    enc = _t(encodings, dtype)
    B, Tv, _ = enc.shape
    r = cfg.step_reduction
    if training:
        teacher = _t(mels, dtype)[:, 0:-1:r, :]
        T = teacher.shape[1]
    else:
        T = cfg.max_step // r
    if steps is not None:
        T = min(T, steps) if training else steps
    u0, u1 = cfg.lstm_sizes
    states = ((torch.zeros(B, u0, dtype=dtype), torch.zeros(B, u0, dtype=dtype)),
              (torch.zeros(B, u1, dtype=dtype), torch.zeros(B, u1, dtype=dtype)))
    align = initial_alignment(cfg, B, Tv, dtype)
    lsa_cum = torch.zeros(B, Tv, dtype=dtype)  # Layers.py:361 sum over all previous alignments
    last = torch.zeros(B, cfg.mel_dim, dtype=dtype)  # Taco2.py:162-165 initial decodings
    values = value_projection(W, enc)
    dec, stops, aligns = [], [], []
    ones0 = torch.ones(B, cfg.prenet_sizes[0], dtype=dtype)
    ones1 = torch.ones(B, cfg.prenet_sizes[1], dtype=dtype)
    zeros_n = torch.zeros(B, Tv, dtype=dtype)
    for t in range(T):
        x = teacher[:, t] if training else last                                     # Taco2.py:183-187
        if reproject_every_step:
            values = value_projection(W, enc)
        k0 = ones0 if keep0 is None else _t(keep0[t], dtype)
        k1 = ones1 if keep1 is None else _t(keep1[t], dtype)
        nz = zeros_n if noise is None else _t(noise[t], dtype)
        mel, stop, align, states, _ = decoder_step(W, cfg, values, x, align, states, k0, k1, nz, lsa_cum)
        lsa_cum = lsa_cum + align
        frames = mel.reshape(B, r, cfg.mel_dim)                                     # Taco2.py:194-201
        last = frames[:, -1]
        dec.append(frames)
        stops.append(stop)
        aligns.append(align)
    return {
        "decodings": torch.cat(dec, dim=1) if dec else torch.zeros(B, 0, cfg.mel_dim, dtype=dtype),
        "stops": torch.cat(stops, dim=1) if stops else torch.zeros(B, 0, dtype=dtype),
        "alignments": torch.stack(aligns, dim=1) if aligns else torch.zeros(B, 0, Tv, dtype=dtype),
        "states": states,
    }


# ----------------------------------------------------------------------------------------
# Postnet (Modules/Taco2.py:130-147 construction, :230 call)
# ----------------------------------------------------------------------------------------
def postnet(weights, cfg, decodings, dtype=torch.float64):
    """post_decodings = Postnet(decodings) + decodings (Taco2.py:230), inference form: per layer a bias-free
    Conv1D(padding='same') (Taco2.py:137-143), BatchNormalization with the moving statistics (:144), tanh on every
    layer but the last two (``index < len(Filters) - 1``, :145-146); the Dropout layers (:147-149) are identity at
    inference.  decodings: [B, T, mel]."""
    W = to_torch(weights, dtype)
    x0 = _t(decodings, dtype)
    x = x0
    for i, (_cout, _k, stride, use_tanh) in enumerate(cfg.postnet_layers):
        x = conv1d_same_nwc(x, W[POST + "/conv1d_%d/kernel" % i], None, stride)
        bn = POST + "/batch_normalization_%d/" % i
        x = batchnorm_inference(x, W[bn + "gamma"], W[bn + "beta"], W[bn + "moving_mean"], W[bn + "moving_variance"])
        if use_tanh:
            x = torch.tanh(x)
    return (x + x0).numpy()


# ----------------------------------------------------------------------------------------
# text Encoder (Modules/Taco2.py:12-51)
# ----------------------------------------------------------------------------------------
def encoder(weights, cfg, tokens, dtype=torch.float64):
    """Encoder.call at inference (Taco2.py:47-51): Embedding (:18-21) -> per conv layer bias-free Conv1D(padding='same')
    (:27-33), BatchNormalization with the moving statistics (:34), ReLU (:35), Dropout = identity (:36-38) ->
    Bidirectional(LSTM(units, return_sequences=True)) (:39-43): forward and time-reversed LSTM from zero states, outputs
    concatenated [forward | backward] per token.  No padding mask anywhere (Embedding has mask_zero=False).
    tokens: [B, T_v] integers -> [B, T_v, 2 * units]."""
    W = to_torch(weights, dtype)
    ids = torch.as_tensor(np.asarray(tokens)).long()
    x = W[ENC + "/embedding/embeddings"][ids]
    for i, stride in enumerate(cfg.encoder_strides):
        x = conv1d_same_nwc(x, W[ENC + "/conv1d_%d/kernel" % i], None, stride)
        bn = ENC + "/batch_normalization_%d/" % i
        x = batchnorm_inference(x, W[bn + "gamma"], W[bn + "beta"], W[bn + "moving_mean"], W[bn + "moving_variance"])
        x = torch.relu(x)
    B, T = x.shape[0], x.shape[1]
    u = cfg.encoder_rnn_size
    outs = []
    for d, order in (("forward_lstm", range(T)), ("backward_lstm", range(T - 1, -1, -1))):
        base = ENC + "/bidirectional/%s/lstm_cell/" % d
        h = torch.zeros(B, u, dtype=dtype)
        c = torch.zeros(B, u, dtype=dtype)
        seq = [None] * T
        for t in order:
            h, c = lstm_cell(x[:, t], h, c, W[base + "kernel"], W[base + "recurrent_kernel"], W[base + "bias"])
            seq[t] = h
        outs.append(torch.stack(seq, dim=1))
    return torch.cat(outs, dim=-1).numpy()


# ----------------------------------------------------------------------------------------
# Vocoder_Taco1 (Modules/Taco2.py:234-260; CBHG :285-385, ConvBank :388-414, Highwaynet :416-434)
# ----------------------------------------------------------------------------------------
def max_pool1d_same(x, pool: int, stride: int):
    """tf.keras.layers.MaxPool1D(padding='same') on [B, W, C]: -inf padding, pad_before = total // 2 (TF 'same' rule)."""
    n = x.shape[1]
    out = -(-n // stride)
    tot = max((out - 1) * stride + pool - n, 0)
    xn = F.pad(x.permute(0, 2, 1), (tot // 2, tot - tot // 2), value=float("-inf"))
    return F.max_pool1d(xn, pool, stride).permute(0, 2, 1)


def vocoder(weights, cfg, mels, dtype=torch.float64, return_parts=False):
    """Vocoder_Taco1.call at inference (Taco2.py:258-260) = Dense(Spectrogram_Dim)(CBHG(mels)).  CBHG.call (:364-376):
    ConvBank = concat over kernel sizes 1..K of ReLU(BatchNormalization(Conv1D('same', no bias))) (:396-413) -> MaxPool1D(pool,
    strides, 'same') (:322-326) -> Conv1D_Projection: [Conv1D('same', no bias) -> BatchNormalization -> ReLU on all but the last]
    then Dense(input channels) when the last filter count differs (:328-345) -> + inputs (:372) -> Highwaynet: Dense(size) when
    the channel count differs, then `count` layers  relu(x W_h + b_h) * s + x * (1 - s),  s = sigmoid(x W_t + b_t) (:347-356,
    430-434) -> Bidirectional(LSTM(rnn_size, return_sequences=True)) (:358-362).  mels: [B, T, Mel_Dim] -> [B, T, Spectrogram_Dim]."""
    W = to_torch(weights, dtype)
    x0 = _t(mels, dtype)

    def bn(x, base):
        return batchnorm_inference(x, W[base + "gamma"], W[base + "beta"], W[base + "moving_mean"], W[base + "moving_variance"])

    bank = []
    for i in range(cfg.voc_bank_count):
        y = conv1d_same_nwc(x0, W[VOC + "/CBHG/ConvBank_%d/conv1d/kernel" % i], None, 1)
        bank.append(torch.relu(bn(y, VOC + "/CBHG/ConvBank_%d/batch_normalization/" % i)))
    x = torch.cat(bank, dim=-1)
    x = max_pool1d_same(x, cfg.voc_pool_size, cfg.voc_pool_strides)
    pooled = x
    n = len(cfg.voc_proj_filters)
    for i in range(n):
        x = conv1d_same_nwc(x, W[VOC + "/CBHG/Conv1D_Projection/conv1d_%d/kernel" % i], None, 1)
        x = bn(x, VOC + "/CBHG/Conv1D_Projection/batch_normalization_%d/" % i)
        if i < n - 1:
            x = torch.relu(x)
    if (VOC + "/CBHG/Conv1D_Projection/dense/kernel") in W:
        x = dense(x, W[VOC + "/CBHG/Conv1D_Projection/dense/kernel"], W[VOC + "/CBHG/Conv1D_Projection/dense/bias"])
    x = x + x0
    if (VOC + "/CBHG/Highwaynet/dense/kernel") in W:
        x = dense(x, W[VOC + "/CBHG/Highwaynet/dense/kernel"], W[VOC + "/CBHG/Highwaynet/dense/bias"])
    for i in range(cfg.voc_highway_count):
        base = VOC + "/CBHG/Highwaynet/highwaynet_%d/" % i
        hgate = torch.relu(dense(x, W[base + "Dense_Relu/kernel"], W[base + "Dense_Relu/bias"]))
        tgate = torch.sigmoid(dense(x, W[base + "Dense_Sigmoid/kernel"], W[base + "Dense_Sigmoid/bias"]))
        x = hgate * tgate + x * (1.0 - tgate)
    highway = x
    B, T = x.shape[0], x.shape[1]
    u = cfg.voc_rnn_size
    outs = []
    for d, order in (("forward_lstm", range(T)), ("backward_lstm", range(T - 1, -1, -1))):
        base = VOC + "/CBHG/RNN/%s/lstm_cell/" % d
        h = torch.zeros(B, u, dtype=dtype)
        c = torch.zeros(B, u, dtype=dtype)
        seq = [None] * T
        for t in order:
            h, c = lstm_cell(x[:, t], h, c, W[base + "kernel"], W[base + "recurrent_kernel"], W[base + "bias"])
            seq[t] = h
        outs.append(torch.stack(seq, dim=1))
    rnn = torch.cat(outs, dim=-1)
    y = dense(rnn, W[VOC + "/Dense/kernel"], W[VOC + "/Dense/bias"])
    if return_parts:
        return dict(pooled=pooled.numpy(), highway=highway.numpy(), rnn=rnn.numpy(), spectrogram=y.numpy())
    return y.numpy()


# ----------------------------------------------------------------------------------------
# GST front end (Modules/GST.py:12-124; Modules/Attention/Layers.py:147-285)
# ----------------------------------------------------------------------------------------
def reference_encoder(W, cfg, mels, mel_lengths):
    """Reference_Encoder.call, GST.py:47-70.  mels [B,T,mel] (initial frame already dropped),
    mel_lengths [B] int."""
    x = mels[:, :, :, None]                                                        # GST.py:54
    for i in range(len(cfg.ref_filters)):
        base = REF + "/Conv2D_{}".format(i)
        x = conv2d_same_nhwc(x, W[base + "/conv2d/kernel"], cfg.ref_strides[i])
        x = batchnorm_inference(x, W[base + "/batch_normalization/gamma"], W[base + "/batch_normalization/beta"],
                                W[base + "/batch_normalization/moving_mean"],
                                W[base + "/batch_normalization/moving_variance"])
        x = torch.relu(x)
    B, Tp = x.shape[0], x.shape[1]
    x = x.reshape(B, Tp, x.shape[2] * x.shape[3])                                  # GST.py:57-62
    seq = gru_sequence(x, W[REF + "/RNN/kernel"], W[REF + "/RNN/recurrent_kernel"], W[REF + "/RNN/bias"])
    lens = torch.as_tensor(np.asarray(mel_lengths)).to(torch.float64)
    idx = torch.ceil(lens / cfg.ref_compress).to(torch.int64) - 1                  # GST.py:38-40,65-68
    picked = seq[torch.arange(B), idx]
    return torch.tanh(dense(picked, W[REF + "/Dense/kernel"], W[REF + "/Dense/bias"]))  # GST.py:70


def layer_norm(x, gamma, beta, eps: float = 1e-8):
    """Layer_Norm.call, Layers.py:280-285: biased variance, eps inside the sqrt."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return gamma * (x - mean) / torch.sqrt(var + eps) + beta


def multi_head_attention(q_kernel, q_bias, v_kernel, v_bias, ln_gamma, ln_beta, num_heads: int,
                         query, value):
    """MultiHeadAttention.call on the 2-input path (key = value), Layers.py:172-214: unscaled dot
    product (use_scale=False), softmax over Tv, heads re-concatenated, Layer_Norm(result + q).
    query [B,Tq,Dq], value [B,Tv,Dv] -> (result [B,Tq,S], head-mean distribution [B,Tq,Tv])."""
    q = dense(query, q_kernel, q_bias)
    v = dense(value, v_kernel, v_bias)
    k = v
    qs = torch.cat(torch.chunk(q, num_heads, dim=-1), dim=0)                       # Layers.py:179-181
    vs = torch.cat(torch.chunk(v, num_heads, dim=-1), dim=0)
    ks = torch.cat(torch.chunk(k, num_heads, dim=-1), dim=0)
    scores = qs @ ks.transpose(1, 2)                                               # Layers.py:224
    dist = torch.softmax(scores, dim=-1)                                           # Layers.py:235
    res = dist @ vs
    res = torch.cat(torch.chunk(res, num_heads, dim=0), dim=-1)                    # Layers.py:209
    res = layer_norm(res + q, ln_gamma, ln_beta)                                   # Layers.py:211
    dist = torch.stack(torch.chunk(dist, num_heads, dim=0), dim=1).mean(dim=1)     # Layers.py:212
    return res, dist


def style_token_layer(weights, cfg, mels_for_gst, mel_lengths, dtype=torch.float64, return_parts=False):
    """Style_Token_Layer.call, GST.py:91-109.  mels_for_gst [B,1+T,mel] (initial frame is dropped,
    GST.py:98), mel_lengths [B].  Returns [B, S]."""
    W = weights if all(isinstance(v, torch.Tensor) and v.dtype == dtype for v in weights.values()) \
        else to_torch(weights, dtype)
    mels = _t(mels_for_gst, dtype)
    ref = reference_encoder(W, cfg, mels[:, 1:], mel_lengths)
    B = ref.shape[0]
    tokens = torch.tanh(W[GST + "/gst_tokens"])[None].expand(B, -1, -1)            # GST.py:100-103
    res, dist = multi_head_attention(
        W[GST + "/Attention/Query/kernel"], W[GST + "/Attention/Query/bias"],
        W[GST + "/Attention/Value/kernel"], W[GST + "/Attention/Value/bias"],
        W[GST + "/Attention/Layer_Normalization/gamma"], W[GST + "/Attention/Layer_Normalization/beta"],
        cfg.style_heads, ref[:, None, :], tokens)
    out = res[:, 0]                                                                # GST.py:109
    if return_parts:
        return out, ref, dist[:, 0]
    return out


def gst_concat(encoders, gsts):
    """GST_Concated_Encoder.call, GST.py:115-124: GST channels FIRST."""
    Tv = encoders.shape[1]
    return torch.cat([gsts[:, None, :].expand(-1, Tv, -1), encoders], dim=-1)


# ----------------------------------------------------------------------------------------
# Counter-based randomness shared with the CUDA path ("philox" RNG mode)
# ----------------------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011), vectorised over numpy uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


STREAM_KEEP0, STREAM_KEEP1, STREAM_NOISE = 0, 1, 2


def _philox_block(seed: int, stream: int, step, row, n_items: int):
    """u32 words for items 0..n_items-1 of (stream, step, row): item i comes from counter
    (i // 4, step, row, stream), word i % 4; key = (seed low, seed high)."""
    nblk = (n_items + 3) // 4
    blk = np.arange(nblk, dtype=np.uint32)
    step = np.asarray(step, dtype=np.uint32)[..., None]
    row = np.asarray(row, dtype=np.uint32)[..., None]
    w = philox4x32_10(blk, step, row, np.uint32(stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.stack(w, axis=-1).reshape(w[0].shape[:-1] + (nblk * 4,))
    return out[..., :n_items]


def philox_keep_mask(seed: int, stream: int, T: int, B: int, n: int, rate: float, t0: int = 0, b0: int = 0):
    """keep iff u >= rate with u = (word >> 8) * 2^-24 (tf.nn.dropout keeps where uniform >= rate).
    Returns float32 [T,B,n] in {0,1}."""
    steps = (np.arange(T, dtype=np.uint32) + np.uint32(t0))[:, None] * np.ones((1, B), np.uint32)
    rows = np.ones((T, 1), np.uint32) * (np.arange(B, dtype=np.uint32) + np.uint32(b0))[None, :]
    w = _philox_block(seed, stream, steps, rows, n)
    u = (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return (u >= np.float32(rate)).astype(np.float32)


def philox_normal(seed: int, T: int, B: int, n: int, t0: int = 0, b0: int = 0):
    """Box-Muller normals: block (w0,w1,w2,w3) -> items 4i..4i+3 =
    (r(w0) cos(2 pi u(w1)), r(w0) sin(2 pi u(w1)), r(w2) cos(2 pi u(w3)), r(w2) sin(2 pi u(w3)))
    with r(w) = sqrt(-2 ln(((w >> 8) + 1) 2^-24)), u(w) = (w >> 8) 2^-24.  float32 [T,B,n]."""
    steps = (np.arange(T, dtype=np.uint32) + np.uint32(t0))[:, None] * np.ones((1, B), np.uint32)
    rows = np.ones((T, 1), np.uint32) * (np.arange(B, dtype=np.uint32) + np.uint32(b0))[None, :]
    n4 = (n + 3) // 4 * 4
    w = _philox_block(seed, STREAM_NOISE, steps, rows, n4).reshape(T, B, n4 // 4, 4)
    u = (w >> np.uint32(8)).astype(np.float64)
    r0 = np.sqrt(-2.0 * np.log((u[..., 0] + 1.0) * 2.0 ** -24))
    r1 = np.sqrt(-2.0 * np.log((u[..., 2] + 1.0) * 2.0 ** -24))
    a0 = 2.0 * math.pi * u[..., 1] * 2.0 ** -24
    a1 = 2.0 * math.pi * u[..., 3] * 2.0 ** -24
    z = np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1)
    return z.reshape(T, B, n4)[..., :n].astype(np.float32)


STREAM_GRIFFIN_LIM = 3


def philox_uniform(seed: int, stream: int, T: int, B: int, n: int, t0: int = 0, b0: int = 0):
    """u = (word >> 8) * 2^-24 in [0, 1): float32 [B, T, n] with step = frame, row = utterance (the initial Griffin-Lim phases
    the CUDA path draws in RNG mode 'philox', csrc/griffin_lim.cuh)."""
    steps = (np.arange(T, dtype=np.uint32) + np.uint32(t0))[:, None] * np.ones((1, B), np.uint32)
    rows = np.ones((T, 1), np.uint32) * (np.arange(B, dtype=np.uint32) + np.uint32(b0))[None, :]
    w = _philox_block(seed, stream, steps, rows, n)
    return np.ascontiguousarray(((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).transpose(1, 0, 2))


def philox_randomness(cfg, seed: int, T: int, B: int, Tv: int, t0: int = 0, b0: int = 0):
    """(keep0 [T,B,p0], keep1 [T,B,p1], noise [T,B,Tv]) exactly as the CUDA path draws them in
    RNG mode 'philox'."""
    k0 = philox_keep_mask(seed, STREAM_KEEP0, T, B, cfg.prenet_sizes[0], cfg.prenet_dropout, t0, b0)
    k1 = philox_keep_mask(seed, STREAM_KEEP1, T, B, cfg.prenet_sizes[1], cfg.prenet_dropout, t0, b0)
    nz = philox_normal(seed, T, B, Tv, t0, b0)
    return k0, k1, nz


# ----------------------------------------------------------------------------------------
# Synthetic inputs
# ----------------------------------------------------------------------------------------
def synth_decoder_inputs(cfg, B: int, Tv: int, T: int, seed: int = 2024, rand_seed: int = 7,
                         teacher: bool = True):
    rng = np.random.default_rng(seed)
    enc = rng.uniform(-1.0, 1.0, size=(B, Tv, cfg.enc_dim)).astype(np.float32)
    mels = None
    if teacher:
        mels = rng.uniform(-4.0, 4.0, size=(B, T * cfg.step_reduction + 1, cfg.mel_dim)).astype(np.float32)
        mels[:, 0] = 0.0  # Feeder.py:125-128 initial zero frame
    rr = np.random.default_rng(rand_seed)
    keep0 = (rr.random((T, B, cfg.prenet_sizes[0])) >= cfg.prenet_dropout).astype(np.float32)
    keep1 = (rr.random((T, B, cfg.prenet_sizes[1])) >= cfg.prenet_dropout).astype(np.float32)
    noise = rr.standard_normal((T, B, Tv)).astype(np.float32)
    return enc, mels, keep0, keep1, noise


def synth_gst_inputs(cfg, B: int, T: int, seed: int = 2024, min_len: Optional[int] = None):
    rng = np.random.default_rng(seed)
    mels = rng.uniform(-4.0, 4.0, size=(B, T + 1, cfg.mel_dim)).astype(np.float32)
    mels[:, 0] = 0.0  # Feeder.py:221-224 zero frame prepended
    lo = max(1, int(0.64 * T)) if min_len is None else min_len
    lengths = rng.integers(lo, T + 1, size=(B,)).astype(np.int32)
    return mels, lengths
