"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/gstk.h declares
(no compute calls: there is no GPU on the CPU test tier)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gstk.h")).read()
    return sorted(set(re.findall(r"\b(gstk_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported(lib_built):
    lib = ctypes.CDLL(lib_built)
    syms = _declared_symbols()
    assert "gstk_decode" in syms and "gstk_gst" in syms and len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.gstk_version() == 1


def test_ctypes_binding_covers_header(lib_built):
    from gst_tacotron_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared_symbols()
    _lib.load()


def test_struct_sizes_match_c(lib_built, tmp_path):
    """sizeof() of the ctypes mirrors == sizeof() seen by a C compiler."""
    from gst_tacotron_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gstk.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(GstkConfig),sizeof(GstkTensorDesc),sizeof(GstkDecodeArgs),sizeof(GstkGstArgs),sizeof(GstkMhaArgs),'
                   'sizeof(GstkPostnetArgs),sizeof(GstkAttentionArgs),sizeof(GstkEncoderArgs),sizeof(GstkVocoderArgs),'
                   'sizeof(GstkGriffinLimArgs),offsetof(GstkDecodeArgs,kernel),offsetof(GstkGriffinLimArgs,spectrogram),sizeof(GstkPrenetArgs));return 0;}')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(_lib.GstkConfig), ctypes.sizeof(_lib.GstkTensorDesc),
                     ctypes.sizeof(_lib.GstkDecodeArgs), ctypes.sizeof(_lib.GstkGstArgs),
                     ctypes.sizeof(_lib.GstkMhaArgs), ctypes.sizeof(_lib.GstkPostnetArgs),
                     ctypes.sizeof(_lib.GstkAttentionArgs), ctypes.sizeof(_lib.GstkEncoderArgs),
                     ctypes.sizeof(_lib.GstkVocoderArgs), ctypes.sizeof(_lib.GstkGriffinLimArgs),
                     _lib.GstkDecodeArgs.kernel.offset, _lib.GstkGriffinLimArgs.spectrogram.offset,
                     ctypes.sizeof(_lib.GstkPrenetArgs)]


def test_no_device_fails_loudly(lib_built):
    """No CPU fallback: without a CUDA device the engine refuses to construct."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.runtime import Engine
    from gst_tacotron_b200._lib import GstkError
    with pytest.raises(GstkError) as ei:
        Engine(load_config())
    assert ei.value.code == 2  # GSTK_ENODEVICE


def test_hparams_reads_reference_schema(tmp_path):
    import json
    from gst_tacotron_b200.hparams import DEFAULT_HP, load_config
    hp = json.loads(json.dumps(DEFAULT_HP))
    hp["Tacotron2"]["Decoder"]["Attention"]["Type"] = "BMA"
    hp["Step_Reduction"] = 2
    hp["GST"]["Style_Token"]["Size"] = 10
    p = tmp_path / "Hyper_Parameters.json"
    p.write_text(json.dumps(hp))
    cfg = load_config(str(p))
    assert cfg.attention_type == "BMA" and cfg.step_reduction == 2 and cfg.n_tokens == 10
    assert cfg.enc_dim == 640 and cfg.proj_dim == 161 and cfg.sigmoid_noise == 0.0 and cfg.ref_compress == 64
    hp["Tacotron2"]["Decoder"]["Attention"]["Type"] = "XYZ"
    p.write_text(json.dumps(hp))
    with pytest.raises(ValueError):
        load_config(str(p))


def test_weight_pack_layout_and_roundtrip(tmp_path):
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.weights import from_named_arrays, init_weights, load_npz, save_npz, weight_spec
    cfg = load_config()
    spec = weight_spec(cfg)
    assert spec["Decoder/Decoder_Step/RNN/cell_0/kernel"] == (384, 4096)
    assert spec["Decoder/Decoder_Step/Projection/kernel"] == (1152, 81)
    assert spec["Style_Token_Layer/Reference_Encoder/RNN/kernel"] == (256, 384)
    assert spec["Style_Token_Layer/Reference_Encoder/RNN/bias"] == (2, 384)
    n_dec = sum(int(np.prod(s)) for k, s in spec.items() if k.startswith("Decoder/"))
    assert n_dec == 14458962
    W = init_weights(cfg)
    save_npz(str(tmp_path / "w.npz"), W)
    W2 = load_npz(str(tmp_path / "w.npz"))
    assert all(np.array_equal(W[k], W2[k]) for k in W)
    renamed = {"model/layer_with_weights-3/" + k + ":0": v for k, v in W.items()}
    W3 = from_named_arrays(cfg, renamed)
    assert all(np.array_equal(W[k], W3[k]) for k in W)


def test_postnet_and_encoder_packs_by_layer_order():
    """weights.postnet_pack / encoder_pack: variables handed over in layer order (Keras auto-names are not stable)."""
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.weights import (ENC, POST, encoder_pack, encoder_spec, init_encoder_weights, init_postnet_weights,
                                           postnet_pack, postnet_spec)
    cfg = load_config()
    WP, WE = init_postnet_weights(cfg), init_encoder_weights(cfg)
    n = len(cfg.postnet_layers)
    bn = lambda base: tuple(base[k] for k in ("gamma", "beta", "moving_mean", "moving_variance"))
    got = postnet_pack(cfg, [WP[POST + "/conv1d_%d/kernel" % i] for i in range(n)],
                       [bn({k: WP[POST + "/batch_normalization_%d/%s" % (i, k)] for k in ("gamma", "beta", "moving_mean", "moving_variance")})
                        for i in range(n)])
    assert list(got) == list(postnet_spec(cfg)) and all(np.array_equal(got[k], WP[k]) for k in got)
    m = len(cfg.encoder_filters)
    cell = lambda d: tuple(WE[ENC + "/bidirectional/%s/lstm_cell/%s" % (d, k)] for k in ("kernel", "recurrent_kernel", "bias"))
    got = encoder_pack(cfg, WE[ENC + "/embedding/embeddings"], [WE[ENC + "/conv1d_%d/kernel" % i] for i in range(m)],
                       [bn({k: WE[ENC + "/batch_normalization_%d/%s" % (i, k)] for k in ("gamma", "beta", "moving_mean", "moving_variance")})
                        for i in range(m)], cell("forward_lstm"), cell("backward_lstm"))
    assert list(got) == list(encoder_spec(cfg)) and all(np.array_equal(got[k], WE[k]) for k in got)
    with pytest.raises(ValueError):
        postnet_pack(cfg, [WP[POST + "/conv1d_0/kernel"]] * (n - 1), [None] * (n - 1))
    with pytest.raises(ValueError):
        encoder_pack(cfg, WE[ENC + "/embedding/embeddings"].T, [WE[ENC + "/conv1d_%d/kernel" % i] for i in range(m)],
                     [bn({k: WE[ENC + "/batch_normalization_%d/%s" % (i, k)] for k in ("gamma", "beta", "moving_mean", "moving_variance")})
                      for i in range(m)], cell("forward_lstm"), cell("backward_lstm"))


def test_encoder_and_postnet_hyper_parameters(tmp_path, monkeypatch):
    """The Encoder / Postnet keys of Hyper_Parameters.json (reference lines 94-108, 122-127) and the vocabulary size taken from
    Token_JSON_Path like Taco2.py:9-10,19."""
    import json
    from gst_tacotron_b200.hparams import config_from_hp
    cfg = config_from_hp({})
    assert (cfg.vocab_size, cfg.encoder_embedding, cfg.encoder_filters, cfg.encoder_kernel, cfg.encoder_rnn_size) == \
        (34, 512, [512, 512, 512], [5, 5, 5], 256)
    assert cfg.postnet_layers == [(512, 5, 1, True)] * 3 + [(512, 5, 1, False), (80, 5, 1, False)]   # tanh: index < len(Filters) - 1
    (tmp_path / "tok.json").write_text(json.dumps({chr(65 + i): i for i in range(21)}))
    monkeypatch.chdir(tmp_path)
    cfg = config_from_hp({"Token_JSON_Path": "tok.json",
                          "Tacotron2": {"Encoder": {"Embedding": {"Size": 128}, "Conv": {"Filters": [64, 64], "Kernel_Size": [3, 7],
                                                                                         "Strides": [1, 1]}, "RNN": {"Size": 64}},
                                        "Decoder": {"Conv": {"Filters": [32, 48], "Kernel_Size": [3, 5], "Strides": [1, 1]}}}})
    assert (cfg.vocab_size, cfg.encoder_embedding, cfg.encoder_filters, cfg.encoder_kernel, cfg.encoder_rnn_size) == \
        (21, 128, [64, 64], [3, 7], 64)
    assert cfg.postnet_layers == [(32, 3, 1, True), (48, 5, 1, False), (80, 5, 1, False)]
    from gst_tacotron_b200.weights import encoder_spec, postnet_spec
    es, ps = encoder_spec(cfg), postnet_spec(cfg)
    assert es["Encoder/embedding/embeddings"] == (21, 128) and es["Encoder/conv1d_1/kernel"] == (7, 64, 64)
    assert es["Encoder/bidirectional/backward_lstm/lstm_cell/recurrent_kernel"] == (64, 256)
    assert ps["Decoder/Postnet/conv1d_2/kernel"] == (5, 48, 80)
