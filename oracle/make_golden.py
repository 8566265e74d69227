"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S UNMODIFIED SOURCES (test infrastructure).

    python oracle/make_golden.py            # all variants -> tests/golden/
    python oracle/make_golden.py --worker NAME OUT   (internal: one variant, run from a scratch CWD)

The reference modules (/root/reference/Modules/...) are imported as they are, on top of the torch-backed
TensorFlow API shim in oracle/tf_shim (TensorFlow itself cannot be installed in this container).  They read
``Hyper_Parameters.json`` from the CWD at import (Modules/Taco2.py:6-10), so every variant runs in its own
scratch directory holding a copy of the reference JSON with the variant's overrides.  Weights come from
gst_tacotron_b200.weights.init_weights(seed) and are written into the reference layers' variables; inputs and
the explicit randomness (dropout keep masks, sigmoid noise) are seeded numpy arrays stored with the outputs.
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

VARIANTS = {
    #  name          : hyper-parameter overrides (dotted keys)
    "sma_r1": {"Tacotron2.Decoder.Attention.Type": "SMA", "Step_Reduction": 1, "Max_Step": 7},
    "bma_r1": {"Tacotron2.Decoder.Attention.Type": "BMA", "Step_Reduction": 1, "Max_Step": 6},
    "sma_r2": {"Tacotron2.Decoder.Attention.Type": "SMA", "Step_Reduction": 2, "Max_Step": 10},
    "gst10": {"GST.Style_Token.Size": 10, "Max_Step": 4},
}


def _set(d, dotted, v):
    ks = dotted.split(".")
    for k in ks[:-1]:
        d = d[k]
    d[ks[-1]] = v


def worker(name, out_path):
    sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), REF, ROOT]
    import numpy as np
    import tensorflow as tf  # the shim
    import torch
    from Modules import GST as RG  # noqa: E402  (reference sources)
    from Modules import Taco2 as RT  # noqa: E402
    from Modules.Attention import Layers as RL  # noqa: E402
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.weights import DEC, GST, POST, REF as REFP, init_postnet_weights, init_weights

    # the shim's default initialisers (glorot_uniform of the layers whose kernels are NOT overwritten from the weight pack:
    # MultiHeadAttention / LocationSensitiveAttention test layers) draw from torch's global generator: seed it, so that a
    # regenerated file is bit-identical to the committed one
    torch.manual_seed(20240607)
    cfg = load_config("Hyper_Parameters.json")
    W = init_weights(cfg, seed=1234, bias_scale=0.05)
    WP = init_postnet_weights(cfg, seed=4321)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    out = {"variant": np.array(name), "weights_seed": np.array(1234), "bias_scale": np.array(0.05)}
    _rng = np.random.default_rng(99)

    class _R(object):  # every drawn array is rounded to float32 so that the stored float32 copy is exact
        def __getattr__(self, k):
            f = getattr(_rng, k)
            return lambda *a, **kw: np.asarray(f(*a, **kw)).astype(np.float32).astype(np.float64)
    rng = _R()

    # ------------------------------------------------------------------ decoder step / loop
    dec = RT.Decoder()
    dec.build(None)
    ds = dec.layer_Dict["Decoder_Step"]
    B, Tv = 2, 11
    enc = rng.uniform(-1, 1, (B, Tv, cfg.enc_dim))
    r = cfg.step_reduction

    def queue_random(k0, k1, nz):
        tf.keras.layers.Dropout.mask_queue[:] = []
        tf.random._normal_queue[:] = []
        for s in range(k0.shape[0]):
            tf.keras.layers.Dropout.mask_queue += [t(k0[s]), t(k1[s])]
            if cfg.sigmoid_noise > 0:
                tf.random._normal_queue.append(t(nz[s])[:, None, :])

    def draw(T):
        k0 = (rng.random((T, B, cfg.prenet_sizes[0])) >= cfg.prenet_dropout).astype(np.float64)
        k1 = (rng.random((T, B, cfg.prenet_sizes[1])) >= cfg.prenet_dropout).astype(np.float64)
        nz = rng.standard_normal((T, B, Tv))
        return k0, k1, nz

    # build variables with one throw-away step, then overwrite them with the weight pack
    k0, k1, nz = draw(1)
    queue_random(k0, k1, nz)
    st0 = ds.get_initial_state(batch_size=B, dtype=tf.float32)
    al0 = ds.get_initial_alignment(B, Tv, tf.float32)
    ds([t(enc), t(np.zeros((B, cfg.mel_dim))), al0, st0], training=False)
    pre = ds.layer_Dict["Prenet"].layer.layers
    pre[0].kernel, pre[0].bias = t(W[DEC + "/Prenet/dense/kernel"]), t(W[DEC + "/Prenet/dense/bias"])
    pre[2].kernel, pre[2].bias = t(W[DEC + "/Prenet/dense_1/kernel"]), t(W[DEC + "/Prenet/dense_1/bias"])
    att = ds.layer_Dict["Attention"]
    att.layer_Dict["Query"].kernel, att.layer_Dict["Query"].bias = t(W[DEC + "/Attention/Query/kernel"]), t(W[DEC + "/Attention/Query/bias"])
    att.layer_Dict["Value"].kernel, att.layer_Dict["Value"].bias = t(W[DEC + "/Attention/Value/kernel"]), t(W[DEC + "/Attention/Value/bias"])
    att.attention_v = t(W[DEC + "/Attention/attention_v"])
    att.attention_score_bias = t(W[DEC + "/Attention/attention_score_bias"])
    for i, cell in enumerate(ds.layer_Dict["RNN"].cells):
        cell.kernel = t(W[DEC + "/RNN/cell_%d/kernel" % i])
        cell.recurrent_kernel = t(W[DEC + "/RNN/cell_%d/recurrent_kernel" % i])
        cell.bias = t(W[DEC + "/RNN/cell_%d/bias" % i])
    ds.layer_Dict["Projection"].kernel, ds.layer_Dict["Projection"].bias = t(W[DEC + "/Projection/kernel"]), t(W[DEC + "/Projection/bias"])

    # (1) one Decoder_Step.call from a non-trivial state
    k0, k1, nz = draw(1)
    queue_random(k0, k1, nz)
    mel_in = rng.uniform(-4, 4, (B, cfg.mel_dim))
    prev_al = rng.random((B, Tv))
    prev_al = (prev_al / prev_al.sum(-1, keepdims=True)).astype(np.float32).astype(np.float64)
    states_in = rng.uniform(-0.5, 0.5, (4, B, cfg.lstm_sizes[0]))
    st = ([t(states_in[0]), t(states_in[1])], [t(states_in[2]), t(states_in[3])])
    m, s, a, ns = ds([t(enc), t(mel_in), t(prev_al), st], training=False)
    out.update(step_enc=enc, step_mel_in=mel_in, step_prev_alignment=prev_al, step_states_in=states_in, step_keep0=k0,
               step_keep1=k1, step_noise=nz, step_mel=m.numpy(), step_stop=s.numpy(), step_alignment=a.numpy(),
               step_states=np.stack([ns[0][0].numpy(), ns[0][1].numpy(), ns[1][0].numpy(), ns[1][1].numpy()]))

    # (2) Decoder.call, teacher forced (training=True): mels include the initial zero frame
    T = 5
    mels = rng.uniform(-4, 4, (B, T * r + 1, cfg.mel_dim))
    mels[:, 0] = 0
    k0, k1, nz = draw(T)
    queue_random(k0, k1, nz)
    d, _, s, a = dec([t(enc), t(mels)], training=True)
    out.update(tf_enc=enc, tf_mels=mels, tf_keep0=k0, tf_keep1=k1, tf_noise=nz, tf_decodings=d.numpy(), tf_stops=s.numpy(),
               tf_alignments=a.numpy())

    # (3) Decoder.call, free running (training=False): Max_Step // r steps from the zero frame (Feeder.py:182-185)
    Tf = cfg.max_step // r
    k0, k1, nz = draw(Tf)
    queue_random(k0, k1, nz)
    # the Postnet (Taco2.py:130-147) was built by that call; give it the Postnet pack (own seed: the draws above and below are
    # the same as before the Postnet was captured) and run Decoder.call again for its second return value (Taco2.py:230)
    layers = dec.layer_Dict["Postnet"].layers
    convs = [l for l in layers if isinstance(l, tf.keras.layers.Conv1D)]
    bns = [l for l in layers if isinstance(l, tf.keras.layers.BatchNormalization)]
    assert len(convs) == len(bns) == len(cfg.postnet_layers)
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        conv.kernel = t(WP[POST + "/conv1d_%d/kernel" % i])
        base = POST + "/batch_normalization_%d/" % i
        bn.gamma, bn.beta = t(WP[base + "gamma"]), t(WP[base + "beta"])
        bn.moving_mean, bn.moving_variance = t(WP[base + "moving_mean"]), t(WP[base + "moving_variance"])
    queue_random(k0, k1, nz)
    d, pd, s, a = dec([t(enc), t(np.zeros((B, 1, cfg.mel_dim)))], training=False)
    out.update(fr_keep0=k0, fr_keep1=k1, fr_noise=nz, fr_decodings=d.numpy(), fr_stops=s.numpy(), fr_alignments=a.numpy(),
               fr_post_decodings=pd.numpy(), postnet_seed=np.array(4321))

    # ------------------------------------------------------------------ GST front end
    stl = RG.Style_Token_Layer()
    Bg, Tg = 3, 150
    gm = rng.uniform(-4, 4, (Bg, Tg + 1, cfg.mel_dim))
    gm[:, 0] = 0
    gl = np.array([150, 64, 65], dtype=np.int32)
    stl([t(gm), torch.as_tensor(gl)])  # builds
    refe = stl.layer_Dict["Reference_Encoder"]
    for i in range(len(cfg.ref_filters)):
        conv, bn, _ = refe.layer_Dict["Conv2D_%d" % i].layers
        base = REFP + "/Conv2D_%d" % i
        conv.kernel = t(W[base + "/conv2d/kernel"])
        bn.gamma, bn.beta = t(W[base + "/batch_normalization/gamma"]), t(W[base + "/batch_normalization/beta"])
        bn.moving_mean, bn.moving_variance = t(W[base + "/batch_normalization/moving_mean"]), t(W[base + "/batch_normalization/moving_variance"])
    g = refe.layer_Dict["RNN"]
    g.kernel, g.recurrent_kernel, g.bias = t(W[REFP + "/RNN/kernel"]), t(W[REFP + "/RNN/recurrent_kernel"]), t(W[REFP + "/RNN/bias"])
    refe.layer_Dict["Dense"].kernel, refe.layer_Dict["Dense"].bias = t(W[REFP + "/Dense/kernel"]), t(W[REFP + "/Dense/bias"])
    mha = stl.layer_Dict["Attention"]
    mha.layer_Dict["Query"].kernel, mha.layer_Dict["Query"].bias = t(W[GST + "/Attention/Query/kernel"]), t(W[GST + "/Attention/Query/bias"])
    mha.layer_Dict["Value"].kernel, mha.layer_Dict["Value"].bias = t(W[GST + "/Attention/Value/kernel"]), t(W[GST + "/Attention/Value/bias"])
    ln = mha.layer_Dict["Layer_Normalization"]
    ln.gamma, ln.beta = t(W[GST + "/Attention/Layer_Normalization/gamma"]), t(W[GST + "/Attention/Layer_Normalization/beta"])
    stl.gst_tokens = t(W[GST + "/gst_tokens"])
    style = stl([t(gm), torch.as_tensor(gl)])
    ref_out = refe([t(gm[:, 1:]), torch.as_tensor(gl)])
    cat = RG.GST_Concated_Encoder()([t(enc[:, :, cfg.style_size:]), t(enc[:, 0, :cfg.style_size])])
    out.update(gst_mels=gm, gst_lengths=gl, gst_style=style.numpy(), gst_ref=ref_out.numpy(), cat_out=cat.numpy())

    # (4) MultiHeadAttention.call on generic inputs (own random variables, stored)
    m2 = RL.MultiHeadAttention(num_heads=8, size=64)
    q_in, v_in = rng.standard_normal((2, 3, 24)), rng.standard_normal((2, 7, 40))
    m2([t(q_in), t(v_in)])
    names = {}
    for nm in ("Query", "Value"):
        names[nm + "_kernel"] = m2.layer_Dict[nm].kernel.numpy()
        m2.layer_Dict[nm].bias = t(rng.standard_normal(64) * 0.1)
        names[nm + "_bias"] = m2.layer_Dict[nm].bias.numpy()
    ln2 = m2.layer_Dict["Layer_Normalization"]
    ln2.gamma, ln2.beta = t(rng.uniform(0.5, 1.5, 64)), t(rng.standard_normal(64) * 0.1)
    res, dist = m2([t(q_in), t(v_in)])
    out.update(mha_q=q_in, mha_v=v_in, mha_out=res.numpy(), mha_dist=dist.numpy(), mha_gamma=ln2.gamma.numpy(), mha_beta=ln2.beta.numpy(),
               **{"mha_" + k: v for k, v in names.items()})

    # (5) sequence-form LocationSensitiveAttention.call: the spec of the step-form LSA extension (Layers.py:289-444)
    if name == "sma_r1":
        lsa = RL.LocationSensitiveAttention(size=16, conv_filters=4, conv_kernel_size=5, conv_stride=1)
        ql, vl = rng.standard_normal((2, 4, 12)), rng.standard_normal((2, 9, 20))
        lsa([t(ql), t(vl)])
        for nm in ("Query", "Value", "Alignment_Dense"):
            lsa.layer_Dict[nm].bias = t(rng.standard_normal(16) * 0.1)
        lsa.layer_Dict["Alignment_Conv"].bias = t(rng.standard_normal(4) * 0.1)
        lsa.bias = t(rng.standard_normal(16) * 0.1)
        ctxs, als = lsa([t(ql), t(vl)])
        out.update(lsa_q=ql, lsa_v=vl, lsa_contexts=ctxs.numpy(), lsa_alignments=als.numpy(), lsa_bias=lsa.bias.numpy(),
                   **{"lsa_%s_%s" % (nm, w): getattr(lsa.layer_Dict[nm], w).numpy()
                      for nm in ("Query", "Value", "Alignment_Dense", "Alignment_Conv") for w in ("kernel", "bias")})
    def pack(v):
        v = np.asarray(v)
        if v.dtype == np.float64 and np.array_equal(v.astype(np.float32).astype(np.float64), v):
            return v.astype(np.float32)  # inputs: exactly representable
        return v                         # outputs stay float64
    np.savez_compressed(out_path, **{k: pack(v) for k, v in out.items()})
    print("wrote", out_path, "%.1f KB" % (os.path.getsize(out_path) / 1024))


def encoder_worker(out_path):
    """The reference's text Encoder (Modules/Taco2.py:12-51) on the shim, default hyper-parameters -> tests/golden/encoder/."""
    sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), REF, ROOT]
    import numpy as np
    import tensorflow as tf  # the shim
    import torch
    from Modules import Taco2 as RT  # noqa: E402  (reference sources)
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.weights import ENC, init_encoder_weights

    cfg = load_config("Hyper_Parameters.json")
    WE = init_encoder_weights(cfg, seed=2468)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    rng = np.random.default_rng(77)
    B, Tv = 3, 13
    tokens = rng.integers(2, cfg.vocab_size, size=(B, Tv)).astype(np.int32)
    tokens[:, 0] = 0          # <S> (Feeder.py:166-174)
    tokens[0, -1] = 1         # <E>
    tokens[1, 9:] = 1         # <E> padding of a shorter sentence (Feeder.py:177-180)
    tokens[2, 5:] = 1
    enc = RT.Encoder()
    enc(tokens, training=False)    # builds the Sequential (Taco2.py:16-45) with the shim's own initial values
    layers = enc.layer.layers
    emb = [l for l in layers if isinstance(l, tf.keras.layers.Embedding)]
    convs = [l for l in layers if isinstance(l, tf.keras.layers.Conv1D)]
    bns = [l for l in layers if isinstance(l, tf.keras.layers.BatchNormalization)]
    bi = [l for l in layers if isinstance(l, tf.keras.layers.Bidirectional)]
    assert len(emb) == 1 and len(bi) == 1 and len(convs) == len(bns) == len(cfg.encoder_filters)
    assert tuple(emb[0].embeddings.shape) == (cfg.vocab_size, cfg.encoder_embedding)
    emb[0].embeddings = t(WE[ENC + "/embedding/embeddings"])
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        assert tuple(conv.kernel.shape) == WE[ENC + "/conv1d_%d/kernel" % i].shape
        conv.kernel = t(WE[ENC + "/conv1d_%d/kernel" % i])
        base = ENC + "/batch_normalization_%d/" % i
        bn.gamma, bn.beta = t(WE[base + "gamma"]), t(WE[base + "beta"])
        bn.moving_mean, bn.moving_variance = t(WE[base + "moving_mean"]), t(WE[base + "moving_variance"])
    for d, lay in (("forward_lstm", bi[0].forward_layer), ("backward_lstm", bi[0].backward_layer)):
        base = ENC + "/bidirectional/%s/lstm_cell/" % d
        for leaf in ("kernel", "recurrent_kernel", "bias"):
            assert tuple(getattr(lay.cell, leaf).shape) == WE[base + leaf].shape
            setattr(lay.cell, leaf, t(WE[base + leaf]))
    y = enc(tokens, training=False)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, tokens=tokens, encodings=y.numpy(), encoder_seed=np.array(2468))
    print("wrote", out_path, "%.1f KB" % (os.path.getsize(out_path) / 1024), tuple(y.shape))


def vocoder_worker(out_path):
    """The reference's Vocoder_Taco1 (Modules/Taco2.py:234-260, CBHG :285-385) on the shim, default hyper-parameters
    -> tests/golden/vocoder/vocoder.npz."""
    sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), REF, ROOT]
    import numpy as np
    import tensorflow as tf  # the shim
    import torch
    from Modules import Taco2 as RT  # noqa: E402  (reference sources)
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.weights import VOC, init_vocoder_weights

    cfg = load_config("Hyper_Parameters.json")
    WV = init_vocoder_weights(cfg, seed=1357)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    rng = np.random.default_rng(99)
    B, T = 2, 21
    mels = (rng.standard_normal((B, T, cfg.mel_dim)) * 1.5).astype(np.float32)
    voc = RT.Vocoder_Taco1()
    voc(t(mels), training=False)    # builds CBHG (Taco2.py:310-362) with the shim's own initial values
    cb = voc.layer_Dict["CBHG"]
    used = set()

    def put(obj, attr, name):
        assert tuple(getattr(obj, attr).shape) == WV[name].shape, (name, tuple(getattr(obj, attr).shape), WV[name].shape)
        setattr(obj, attr, t(WV[name]))
        used.add(name)

    def put_bn(bn, base):
        for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            put(bn, leaf, base + leaf)

    for i in range(cfg.voc_bank_count):
        conv, bn, relu = cb.layer_Dict["ConvBank"].layer_Dict["ConvBank_%d" % i].layers
        assert isinstance(conv, tf.keras.layers.Conv1D) and isinstance(bn, tf.keras.layers.BatchNormalization)
        put(conv, "kernel", VOC + "/CBHG/ConvBank_%d/conv1d/kernel" % i)
        put_bn(bn, VOC + "/CBHG/ConvBank_%d/batch_normalization/" % i)
    proj = cb.layer_Dict["Conv1D_Projection"].layers
    convs = [l for l in proj if isinstance(l, tf.keras.layers.Conv1D)]
    bns = [l for l in proj if isinstance(l, tf.keras.layers.BatchNormalization)]
    dns = [l for l in proj if isinstance(l, tf.keras.layers.Dense)]
    assert len(convs) == len(bns) == len(cfg.voc_proj_filters) and len(dns) == 1
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        put(conv, "kernel", VOC + "/CBHG/Conv1D_Projection/conv1d_%d/kernel" % i)
        put_bn(bn, VOC + "/CBHG/Conv1D_Projection/batch_normalization_%d/" % i)
    put(dns[0], "kernel", VOC + "/CBHG/Conv1D_Projection/dense/kernel")
    put(dns[0], "bias", VOC + "/CBHG/Conv1D_Projection/dense/bias")
    hw = cb.layer_Dict["Highwaynet"].layers
    assert isinstance(hw[0], tf.keras.layers.Dense) and len(hw) == 1 + cfg.voc_highway_count
    put(hw[0], "kernel", VOC + "/CBHG/Highwaynet/dense/kernel")
    put(hw[0], "bias", VOC + "/CBHG/Highwaynet/dense/bias")
    for i, lay in enumerate(hw[1:]):
        for nm in ("Dense_Relu", "Dense_Sigmoid"):
            lay.layer_Dict[nm]._maybe_build(torch.zeros(1, 1, cfg.voc_highway_size, dtype=torch.float64))
            put(lay.layer_Dict[nm], "kernel", VOC + "/CBHG/Highwaynet/highwaynet_%d/%s/kernel" % (i, nm))
            put(lay.layer_Dict[nm], "bias", VOC + "/CBHG/Highwaynet/highwaynet_%d/%s/bias" % (i, nm))
    bi = cb.layer_Dict["RNN"]
    for d, lay in (("forward_lstm", bi.forward_layer), ("backward_lstm", bi.backward_layer)):
        for leaf in ("kernel", "recurrent_kernel", "bias"):
            put(lay.cell, leaf, VOC + "/CBHG/RNN/%s/lstm_cell/%s" % (d, leaf))
    put(voc.layer_Dict["Dense"], "kernel", VOC + "/Dense/kernel")
    put(voc.layer_Dict["Dense"], "bias", VOC + "/Dense/bias")
    assert used == set(WV), sorted(set(WV) - used)
    y = voc(t(mels), training=False)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, mels=mels, spectrogram=y.numpy(), vocoder_seed=np.array(1357))
    print("wrote", out_path, "%.1f KB" % (os.path.getsize(out_path) / 1024), tuple(y.shape))


def audio_worker(out_path):
    """The reference's Audio.inv_spectrogram (Audio.py:23-27, 57-68) executed from its own source.  librosa is not installed:
    a stand-in module supplies stft / istft from torch (center=True, reflect padding, periodic Hann window - the convention
    librosa documents), so the golden pins everything the reference itself wrote around them (de-normalisation, dB -> amplitude,
    the power, the Griffin-Lim loop, the inverse pre-emphasis); np.random.rand is replaced by a recorded array."""
    import types
    import numpy as np
    import torch

    def _stft(y, n_fft, hop_length, win_length):
        assert win_length == n_fft
        w = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
        return torch.stft(torch.as_tensor(np.asarray(y, np.float64)), n_fft, hop_length, win_length, w, center=True,
                          pad_mode="reflect", return_complex=True).numpy()

    def _istft(D, hop_length, win_length):
        n_fft = 2 * (D.shape[0] - 1)
        w = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
        return torch.istft(torch.as_tensor(np.asarray(D, np.complex128)), n_fft, hop_length, win_length, w, center=True).numpy()

    lib = types.ModuleType("librosa")
    lib.stft = lambda y, n_fft, hop_length, win_length: _stft(y, n_fft, hop_length, win_length)
    lib.istft = lambda y, hop_length, win_length: _istft(y, hop_length, win_length)
    lib.filters = types.ModuleType("librosa.filters")
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = lib.filters
    if not hasattr(np, "complex"):
        np.complex = complex   # numpy < 1.24 spelling used at Audio.py:62
    sys.path[:0] = [REF]
    import Audio as RA  # noqa: E402  (reference source)
    rng = np.random.default_rng(4242)
    out = {}
    for tag, (F_, T, mav, iters) in {"a": (513, 9, 4, 3), "b": (513, 14, None, 5), "c": (65, 12, 4, 60)}.items():
        spec = (rng.uniform(-1.2, 1.2, (F_, T)) * (mav or 1.0)).astype(np.float32) if mav else rng.uniform(-0.1, 1.1, (F_, T)).astype(np.float32)
        u = rng.random((F_, T))
        real_rand = np.random.rand
        np.random.rand = lambda *shape: u.reshape(shape)
        try:
            wav = RA.inv_spectrogram(spectrogram=spec, num_freq=F_, hop_length=(F_ - 1) // 2, win_length=(F_ - 1) * 2,
                                     sample_rate=16000, max_abs_value=mav, griffin_lim_iters=iters)
        finally:
            np.random.rand = real_rand
        out.update({tag + "_spec": spec, tag + "_uniform": u, tag + "_wav": np.asarray(wav, np.float64),
                    tag + "_max_abs": np.array(-1.0 if mav is None else float(mav)), tag + "_iters": np.array(iters)})
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "%.1f KB" % (os.path.getsize(out_path) / 1024))


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--vocoder-worker":
        return vocoder_worker(sys.argv[2])
    if len(sys.argv) >= 3 and sys.argv[1] == "--audio-worker":
        return audio_worker(sys.argv[2])
    if len(sys.argv) >= 2 and sys.argv[1] in ("--vocoder", "--audio"):   # only these goldens (the others stay untouched)
        hp0 = json.load(open(os.path.join(REF, "Hyper_Parameters.json")))
        which = sys.argv[1][2:]
        with tempfile.TemporaryDirectory() as td:
            json.dump(hp0, open(os.path.join(td, "Hyper_Parameters.json"), "w"))
            json.dump(json.load(open(os.path.join(REF, hp0["Token_JSON_Path"]))), open(os.path.join(td, hp0["Token_JSON_Path"]), "w"))
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--%s-worker" % which,
                                   os.path.join(ROOT, "tests", "golden", "vocoder", which + ".npz")], cwd=td)
        return
    if len(sys.argv) >= 4 and sys.argv[1] == "--worker":
        return worker(sys.argv[2], sys.argv[3])
    if len(sys.argv) >= 3 and sys.argv[1] == "--encoder-worker":
        return encoder_worker(sys.argv[2])
    if len(sys.argv) >= 2 and sys.argv[1] == "--encoder":   # only the Encoder golden (the decoder/GST goldens stay untouched)
        hp0 = json.load(open(os.path.join(REF, "Hyper_Parameters.json")))
        with tempfile.TemporaryDirectory() as td:
            json.dump(hp0, open(os.path.join(td, "Hyper_Parameters.json"), "w"))
            json.dump(json.load(open(os.path.join(REF, hp0["Token_JSON_Path"]))), open(os.path.join(td, hp0["Token_JSON_Path"]), "w"))
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--encoder-worker",
                                   os.path.join(ROOT, "tests", "golden", "encoder", "encoder.npz")], cwd=td)
        return
    hp0 = json.load(open(os.path.join(REF, "Hyper_Parameters.json")))
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, over in VARIANTS.items():
        hp = json.loads(json.dumps(hp0))
        for k, v in over.items():
            _set(hp, k, v)
        with tempfile.TemporaryDirectory() as td:
            json.dump(hp, open(os.path.join(td, "Hyper_Parameters.json"), "w"))
            json.dump(json.load(open(os.path.join(REF, hp0["Token_JSON_Path"]))), open(os.path.join(td, hp0["Token_JSON_Path"]), "w"))
            json.dump(over, open(os.path.join(ROOT, "tests", "golden", name + ".hp.json"), "w"))
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--worker", name,
                                   os.path.join(ROOT, "tests", "golden", name + ".npz")], cwd=td)


if __name__ == "__main__":
    main()
