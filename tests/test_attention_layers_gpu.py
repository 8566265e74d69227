"""Stand-alone attention drop-ins (Steps.py / Layers.py signatures) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests.util import FP32_TOL, make_cfg, make_weights, max_abs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA")
    e = Engine(cfg, make_weights(cfg))
    yield e
    e.close()


def _ref(layer_type, w, q, v, k, prev, noise, sn):
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    W = {O.DEC + "/Attention/attention_v": t(w["attention_v"]), O.DEC + "/Attention/attention_score_bias": t(w["attention_score_bias"])}
    qp = t(q) @ t(w["Query/kernel"]) + t(w["Query/bias"])
    vp = t(v) @ t(w["Value/kernel"]) + t(w["Value/bias"])
    kp = vp if k is None else t(k) @ t(w["Key/kernel"]) + t(w["Key/bias"])
    e = O.monotonic_scores(W, qp, kp)
    fn = O.sma_probability if layer_type == "SMA" else O.bma_probability
    al = fn(e, t(prev), t(noise) if noise is not None else torch.zeros_like(e), sn if noise is not None else 0.0)
    return torch.einsum("bt,bta->ba", al, vp).numpy(), al.numpy()


@pytest.mark.parametrize("kind", ["SMA", "BMA"])
@pytest.mark.parametrize("four_inputs", [False, True])
def test_step_attention_layers(eng, kind, four_inputs):
    from gst_tacotron_b200.Modules.Attention.Steps import BahdanauMonotonicAttention, StepwiseMonotonicAttention
    rng = np.random.default_rng(3)
    B, Tv, Dq, Dv, Dk, A = 3, 41, 24, 40, 20, 64
    q = rng.standard_normal((B, Dq)).astype(np.float32)
    v = rng.standard_normal((B, Tv, Dv)).astype(np.float32)
    k = rng.standard_normal((B, Tv, Dk)).astype(np.float32) if four_inputs else None
    prev = rng.random((B, Tv)).astype(np.float32)
    prev /= prev.sum(-1, keepdims=True)
    noise = rng.standard_normal((B, Tv)).astype(np.float32)
    layer = (StepwiseMonotonicAttention if kind == "SMA" else BahdanauMonotonicAttention)(A, engine=eng, seed=5)
    assert layer.sigmoid_noise == (2.0 if kind == "SMA" else 0.0)  # Steps.py:58,212
    inputs = [q, v, k, prev] if four_inputs else [q, v, prev]
    nz = noise if kind == "SMA" else None
    ctx, al = layer(inputs, noise=nz)
    layer.weights["attention_score_bias"] = np.float32(0.3) * np.ones((), np.float32)
    layer.weights["Query/bias"] = rng.standard_normal(A).astype(np.float32) * 0.1
    ctx, al = layer(inputs, noise=nz)
    rctx, ral = _ref(kind, layer.weights, q, v, k, prev, nz, layer.sigmoid_noise)
    assert max_abs(ctx, rctx) < FP32_TOL and max_abs(al, ral) < FP32_TOL
    a0 = layer.initial_alignment_fn(B, Tv)
    assert a0.shape == (B, Tv) and float(a0[:, 0].sum()) == B and float(a0.sum()) == B
    with pytest.raises(ValueError):
        layer([q, v])


def test_multi_head_attention_layer(eng):
    from gst_tacotron_b200.Modules.Attention.Layers import MultiHeadAttention
    with pytest.raises(ValueError):
        MultiHeadAttention(num_heads=5, size=64)
    rng = np.random.default_rng(4)
    q = rng.standard_normal((2, 3, 24)).astype(np.float32)
    v = rng.standard_normal((2, 9, 40)).astype(np.float32)
    m = MultiHeadAttention(num_heads=4, size=64, engine=eng)
    out, dist = m([q, v])
    w = m.weights
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    ref, rdist = O.multi_head_attention(t(w["Query/kernel"]), t(w["Query/bias"]), t(w["Value/kernel"]), t(w["Value/bias"]),
                                        t(w["Layer_Normalization/gamma"]), t(w["Layer_Normalization/beta"]), 4, t(q), t(v))
    assert max_abs(out, ref) < FP32_TOL and max_abs(dist, rdist) < FP32_TOL


def test_prenet_layer_stand_alone():
    """Modules.Taco2.Prenet(sizes, dropout_rate)(inputs, training) (Taco2.py:262-283; dropout always on): external masks against the
    oracle, the Philox mode against the oracle fed the decoder's own streams, rng='none' = plain Dense(relu) x 2."""
    import torch
    from gst_tacotron_b200.Modules.Taco2 import Prenet
    from gst_tacotron_b200.runtime import Engine
    from oracle import reference_port as O
    cfg = make_cfg("SMA")
    W = make_weights(cfg)
    eng = Engine(cfg, W)
    try:
        rng = np.random.default_rng(0)
        B, T = 5, 7
        x = rng.uniform(-4, 4, (B, T, cfg.mel_dim)).astype(np.float32)
        k0 = (rng.random((B, T, cfg.prenet_sizes[0])) >= cfg.prenet_dropout).astype(np.float32)
        k1 = (rng.random((B, T, cfg.prenet_sizes[1])) >= cfg.prenet_dropout).astype(np.float32)
        Wt = O.to_torch(W)
        t64 = lambda a: torch.as_tensor(a, dtype=torch.float64)
        layer = Prenet(cfg.prenet_sizes, cfg.prenet_dropout, engine=eng)
        out = layer(x, training=False, rng="external", keep0=k0, keep1=k1)     # `training` is ignored: always on (Taco2.py:283)
        assert out.shape == (B, T, cfg.prenet_sizes[1])
        assert max_abs(out, O.prenet(Wt, cfg, t64(x), t64(k0), t64(k1)).numpy()) < FP32_TOL
        plain = layer(torch.as_tensor(x, device="cuda:0"), rng="none")
        h0 = torch.relu(O.dense(t64(x), Wt[O.DEC + "/Prenet/dense/kernel"], Wt[O.DEC + "/Prenet/dense/bias"]))
        ref_plain = torch.relu(O.dense(h0, Wt[O.DEC + "/Prenet/dense_1/kernel"], Wt[O.DEC + "/Prenet/dense_1/bias"])).numpy()
        assert plain.is_cuda and max_abs(plain, ref_plain) < FP32_TOL
        # Philox: rows = utterances of decoder step 3 (the masks Decoder_Step would draw there)
        xs = x[:, 0]
        pk0, pk1, _ = O.philox_randomness(cfg, 9, 4, B, 8)
        got = layer(xs, rng="philox", seed=9, step=3)
        assert max_abs(got, O.prenet(Wt, cfg, t64(xs), t64(pk0[3]), t64(pk1[3])).numpy()) < FP32_TOL
        with pytest.raises(ValueError):
            Prenet([128, 128], 0.5, engine=eng)(xs)
    finally:
        eng.close()
