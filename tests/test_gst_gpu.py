"""GPU parity of the GST front end (Reference_Encoder, Style_Token_Layer, MultiHeadAttention)."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests.util import FP32_TOL, make_cfg, make_weights, max_abs, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng16():
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA")
    W = make_weights(cfg)
    e = Engine(cfg, W)
    yield cfg, W, e
    e.close()


@pytest.mark.parametrize("B,T", [(1, 188), (4, 257), (3, 64), (5, 1000)])
def test_style_token_layer_matches_oracle(eng16, B, T):
    cfg, W, eng = eng16
    mels, lens = O.synth_gst_inputs(cfg, B, T, min_len=1)
    ref, ref_enc, ref_att = O.style_token_layer(W, cfg, mels, lens, return_parts=True)
    out = eng.gst(mels, lens, drop_first=True, want=("gst", "ref", "attention"))
    assert max_abs(out["ref"], ref_enc) < FP32_TOL
    assert max_abs(out["gst"], ref) < FP32_TOL   # measured 1e-6 (tools/gst_layernorm_gain.py: Layer_Norm gain <= 6.5 on 2e-7)
    assert max_abs(out["attention"], ref_att) < FP32_TOL
    assert np.allclose(to_np(out["attention"]).sum(-1), 1.0, atol=1e-5)
    g = to_np(out["gst"])
    assert np.allclose(g.mean(-1), 0.0, atol=1e-4)  # Layer_Norm with gamma=1, beta=0


def test_ten_tokens_variant():
    """BASELINE config 4 names a 10-token bank; the shipped default is 16 (Hyper_Parameters.json:28-29)."""
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA", n_tokens=10)
    W = make_weights(cfg)
    eng = Engine(cfg, W)
    mels, lens = O.synth_gst_inputs(cfg, 6, 333)
    ref = O.style_token_layer(W, cfg, mels, lens)
    out = eng.gst(torch.as_tensor(mels, device="cuda:0"), lens, want=("gst",))
    assert out["gst"].is_cuda
    assert max_abs(out["gst"], ref) < FP32_TOL
    eng.close()


def test_reference_encoder_layer_and_lengths(eng16):
    cfg, W, eng = eng16
    from gst_tacotron_b200.Modules.GST import Reference_Encoder, Style_Token_Layer
    B, T = 4, 200
    mels, _ = O.synth_gst_inputs(cfg, B, T)
    lens = np.array([1, 64, 65, 200], np.int32)  # gather index ceil(len/64)-1 = 0,0,1,3
    Wt = O.to_torch(W)
    ref = O.reference_encoder(Wt, cfg, torch.as_tensor(mels[:, 1:], dtype=torch.float64), lens)
    out = Reference_Encoder(eng)([mels[:, 1:].copy(), lens])
    assert max_abs(out, ref) < FP32_TOL
    out2 = Style_Token_Layer(eng)([mels, lens])
    assert max_abs(out2, O.style_token_layer(W, cfg, mels, lens)) < FP32_TOL


def test_generic_multi_head_attention(eng16):
    cfg, W, eng = eng16
    rng = np.random.default_rng(5)
    B, tq, tv, dq, dv, S, H = 3, 2, 7, 24, 40, 64, 8
    q = rng.standard_normal((B, tq, dq)).astype(np.float32)
    v = rng.standard_normal((B, tv, dv)).astype(np.float32)
    Wq = (rng.standard_normal((dq, S)) * 0.2).astype(np.float32)
    Wv = (rng.standard_normal((dv, S)) * 0.2).astype(np.float32)
    bq = rng.standard_normal(S).astype(np.float32) * 0.1
    bv = rng.standard_normal(S).astype(np.float32) * 0.1
    g = rng.uniform(0.5, 1.5, S).astype(np.float32)
    b = rng.standard_normal(S).astype(np.float32) * 0.1
    t = lambda x: torch.as_tensor(x, dtype=torch.float64)
    ref, ref_att = O.multi_head_attention(t(Wq), t(bq), t(Wv), t(bv), t(g), t(b), H, t(q), t(v))
    out, att = eng.mha(q, v, Wq, bq, Wv, bv, g, b, H)
    assert max_abs(out, ref) < FP32_TOL
    assert max_abs(att, ref_att) < FP32_TOL
    with pytest.raises(ValueError):
        eng.mha(q, v, Wq, bq, Wv, bv, g, b, 5)  # size % heads != 0 (Layers.py:155-156)


# ---- tensor-core mode (handle precision "bf16"): conv stack as implicit GEMMs on tcgen05 (csrc/gst_tc.cuh), fp16 operands.
# Tolerance: 1e-2 absolute (north_star, bf16 mode) on the Reference_Encoder output, the token weights and the style embedding.
@pytest.fixture(scope="module")
def eng_tc():
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA", precision="bf16")
    W = make_weights(cfg)
    e = Engine(cfg, W)
    yield cfg, W, e
    e.close()


# frame counts with odd / even sizes at different depths of the stride-2 stack (TF 'same': one padding row in front of odd sizes)
@pytest.mark.parametrize("B,T", [(1, 188), (4, 257), (3, 64), (2, 65), (5, 1000), (7, 999), (130, 131)])
def test_tensor_core_conv_stack_matches_oracle(eng_tc, B, T):
    cfg, W, eng = eng_tc
    mels, lens = O.synth_gst_inputs(cfg, B, T, min_len=1)
    ref, ref_enc, ref_att = O.style_token_layer(W, cfg, mels, lens, return_parts=True)
    out = eng.gst(mels, lens, drop_first=True, want=("gst", "ref", "attention"))
    assert max_abs(out["ref"], ref_enc) < 1e-2
    assert max_abs(out["attention"], ref_att) < 1e-2
    assert max_abs(out["gst"], ref) < 1e-2
    assert np.allclose(to_np(out["attention"]).sum(-1), 1.0, atol=1e-5)


def test_tensor_core_conv_stack_agrees_with_ffma_kernels(eng_tc, monkeypatch):
    """same handle, GSTK_GST_TC=0 -> fp32 FFMA convolutions: the two paths differ by fp16 operand rounding only"""
    cfg, W, eng = eng_tc
    mels, lens = O.synth_gst_inputs(cfg, 6, 333)
    a = eng.gst(mels, lens, want=("gst", "ref"))
    monkeypatch.setenv("GSTK_GST_TC", "0")
    b = eng.gst(mels, lens, want=("gst", "ref"))
    assert 0 < max_abs(a["ref"], b["ref"]) < 1e-2 and max_abs(a["gst"], b["gst"]) < 1e-2
    ref = O.style_token_layer(W, cfg, mels, lens)
    assert max_abs(b["gst"], ref) < FP32_TOL


def test_tensor_core_conv_stack_other_filters():
    """layer counts / channel counts off the defaults (the library accepts 32/64/128/256 filters)"""
    from gst_tacotron_b200.runtime import Engine
    for filters in ([32, 64, 128], [64, 32], [256, 32, 32, 64]):
        cfg = make_cfg("SMA", precision="bf16", ref_filters=filters, ref_kernel=[3] * len(filters), ref_strides=[2] * len(filters))
        W = make_weights(cfg)
        eng = Engine(cfg, W)
        try:
            mels, lens = O.synth_gst_inputs(cfg, 3, 120)
            ref, ref_enc, _ = O.style_token_layer(W, cfg, mels, lens, return_parts=True)
            out = eng.gst(mels, lens, want=("gst", "ref"))
            assert max_abs(out["ref"], ref_enc) < 1e-2 and max_abs(out["gst"], ref) < 1e-2
        finally:
            eng.close()
