"""GPU parity: persistent CUDA decoder (through the C ABI) vs the CPU oracle on the same seeded
inputs.  Tolerances are the ones BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests.util import FP32_TOL, make_cfg, make_weights, max_abs, oracle_decode, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from gst_tacotron_b200.runtime import Engine
    cache = {}

    def get(attention="SMA", **over):
        key = (attention, tuple(sorted(over.items())))
        if key not in cache:
            cfg = make_cfg(attention, **over)
            W = make_weights(cfg)
            cache[key] = (cfg, W, Engine(cfg, W))
        return cache[key]
    yield get
    for _, _, e in cache.values():
        e.close()


def _check(out, ref, T, tol=FP32_TOL):
    assert max_abs(out["mel"], ref["decodings"]) < tol
    assert max_abs(out["stop"], ref["stops"]) < tol
    assert max_abs(out["alignment"], ref["alignments"]) < tol


@pytest.mark.parametrize("attention", ["SMA", "BMA", "LSA"])
@pytest.mark.parametrize("B,Tv,T", [(1, 82, 12), (3, 37, 20), (17, 50, 6)])
def test_teacher_forced_matches_oracle(engines, attention, B, Tv, T):
    cfg, W, eng = engines(attention)
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, 0:-1:cfg.step_reduction], rng="external",
                     keep0=k0, keep1=k1, noise=nz)
    _check(out, ref, T)


@pytest.mark.parametrize("attention", ["SMA", "BMA", "LSA"])
def test_free_running_matches_oracle(engines, attention):
    cfg, W, eng = engines(attention)
    B, Tv, T = 2, 40, 25
    enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
    ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
    _check(out, ref, T, tol=1e-3)  # feedback amplifies rounding; north_star asks for stop-frame identity
    assert np.array_equal(to_np(out["stop"]) < 0, ref["stops"] < 0)


def test_no_randomness_mode(engines):
    cfg, W, eng = engines("SMA")
    B, Tv, T = 2, 33, 10
    enc, mels, _, _, _ = O.synth_decoder_inputs(cfg, B, Tv, T)
    import copy
    cfg0 = copy.deepcopy(cfg)
    cfg0.prenet_dropout = 0.0
    ref = O.decoder_loop(W, cfg0, enc, mels=mels, training=True, noise=None)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="none")
    assert max_abs(out["mel"], ref["decodings"]) < FP32_TOL
    assert max_abs(out["alignment"], ref["alignments"]) < FP32_TOL


def test_philox_mode_matches_oracle_streams(engines):
    cfg, W, eng = engines("SMA")
    B, Tv, T = 5, 45, 9
    enc, mels, _, _, _ = O.synth_decoder_inputs(cfg, B, Tv, T)
    k0, k1, nz = O.philox_randomness(cfg, seed=0x1234ABCD5678, T=T, B=B, Tv=Tv, t0=3, b0=11)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="philox", seed=0x1234ABCD5678, step_offset=3,
                     row_offset=11)
    _check(out, ref, T)


def test_gst_concat_folded_into_value_projection(engines):
    cfg, W, eng = engines("SMA")
    B, Tv, T = 3, 29, 5
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    gst = enc[:, 0, :cfg.style_size].copy()
    text = enc[:, :, cfg.style_size:].copy()
    full = np.concatenate([np.repeat(gst[:, None, :], Tv, axis=1), text], axis=-1)  # GST.py:121-124
    a = eng.decode(encodings=full, teacher_mels=mels[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz)
    b = eng.decode(enc_text=text, gst=gst, teacher_mels=mels[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz)
    assert max_abs(a["mel"], b["mel"]) < 1e-5
    cat = eng.concat_encoder(text, gst)
    assert np.array_equal(to_np(cat), full)


def test_more_utterances_than_sms(engines):
    cfg, W, eng = engines("SMA")
    B, Tv, T = 160, 21, 3
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz)
    _check(out, ref, T)


def test_device_tensors_and_states_roundtrip(engines):
    """Splitting a decode into two calls through init_* state must equal one call (Decoder_Step contract)."""
    cfg, W, eng = engines("SMA")
    B, Tv, T = 4, 31, 8
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    dev = "cuda:0"
    enc_d = torch.as_tensor(enc, device=dev)
    teach = torch.as_tensor(mels[:, :-1], device=dev)
    full = eng.decode(encodings=enc_d, teacher_mels=teach, rng="philox", seed=9)
    assert isinstance(full["mel"], torch.Tensor) and full["mel"].is_cuda
    h = 5
    a = eng.decode(encodings=enc_d, teacher_mels=teach[:, :h].contiguous(), rng="philox", seed=9,
                   want=("mel", "stop", "alignment", "states"))
    b = eng.decode(encodings=enc_d, teacher_mels=teach[:, h:].contiguous(), rng="philox", seed=9, step_offset=h,
                   init_alignment=a["alignment"][:, -1].contiguous(), init_states=a["states"])
    assert max_abs(torch.cat([a["mel"], b["mel"]], 1), full["mel"]) < 1e-6
    assert max_abs(torch.cat([a["alignment"], b["alignment"]], 1), full["alignment"]) < 1e-6


def test_step_reduction_2(engines):
    cfg, W, eng = engines("SMA", step_reduction=2)
    B, Tv, T = 2, 30, 6
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, 0:-1:2], rng="external", keep0=k0, keep1=k1, noise=nz)
    assert ref["decodings"].shape == (B, T * 2, cfg.mel_dim)
    _check(out, ref, T)
    reff = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
    outf = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
    _check(outf, reff, T, tol=1e-3)


def test_sma_alignment_properties(engines):
    """Analytic checks that follow from Steps.py:201-229: rows start one-hot, stay non-negative, and
    mass only leaks off the right edge (sum <= 1, == 1 while the edge is unreachable)."""
    cfg, W, eng = engines("SMA")
    B, Tv, T = 3, 60, 40
    enc, _, _, _, _ = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
    out = eng.decode(encodings=enc, steps=T, rng="philox", seed=1)
    al = to_np(out["alignment"])
    assert al.min() >= 0.0
    s = al.sum(-1)
    assert np.all(s <= 1.0 + 1e-5)
    assert np.allclose(s[:, : Tv - 1], 1.0, atol=1e-5)
    # support can move at most one position per step
    assert np.all(al[:, 0, 2:] == 0.0)


def test_config2_teacher_forced_full_size(engines):
    """BASELINE config 2: batch 64, 150 tokens x 800 frames.  Full size on the GPU; the oracle checks a
    200-step prefix (teacher forcing makes prefixes independent of later frames)."""
    cfg, W, eng = engines("SMA")
    B, Tv, T = 64, 150, 800
    enc, mels, _, _, _ = O.synth_decoder_inputs(cfg, B, Tv, T)
    seed = 77
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="philox", seed=seed)
    Tp = 120
    k0, k1, nz = O.philox_randomness(cfg, seed, Tp, B, Tv)
    ref = oracle_decode(cfg, W, enc, mels=mels[:, : Tp + 1], keep0=k0, keep1=k1, noise=nz, dtype=torch.float32)
    assert max_abs(to_np(out["mel"])[:, :Tp], ref["decodings"]) < FP32_TOL
    assert max_abs(to_np(out["stop"])[:, :Tp], ref["stops"]) < FP32_TOL
    assert max_abs(to_np(out["alignment"])[:, :Tp], ref["alignments"]) < FP32_TOL
    al = to_np(out["alignment"])
    assert np.isfinite(to_np(out["mel"])).all() and al.min() >= 0 and np.all(al.sum(-1) <= 1 + 1e-4)


def test_bad_arguments_raise(engines):
    cfg, W, eng = engines("SMA")
    enc = np.zeros((1, 4, cfg.enc_dim), np.float32)
    with pytest.raises(ValueError):
        eng.decode(encodings=enc, steps=2, rng="external")  # masks missing
    from gst_tacotron_b200.runtime import Engine
    import copy
    bad = copy.deepcopy(cfg)
    bad.attention_type = "XYZ"
    with pytest.raises(ValueError):
        Engine(bad, W)


def _bf16_round(x):
    import torch as _t
    return _t.as_tensor(x).to(_t.bfloat16).to(_t.float32).numpy()


@pytest.mark.parametrize("K", [64, 128, 512])
def test_tcgen05_tile_selftest(engines, K):
    """One UMMA tile through the decoder's operand packing / descriptors / TMEM loads vs numpy."""
    import ctypes as C
    cfg, W, eng = engines("SMA")
    rng = np.random.default_rng(K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    Bm = rng.standard_normal((32, K)).astype(np.float32)
    D = np.zeros((128, 32), np.float32)
    rc = eng._lib.gstk_selftest_umma(eng._h, A.ctypes.data, Bm.ctypes.data, K, D.ctypes.data)
    assert rc == 0, eng._lib.gstk_last_error(eng._h)
    ref = _bf16_round(A).astype(np.float64) @ _bf16_round(Bm).astype(np.float64).T
    assert np.abs(D - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())
