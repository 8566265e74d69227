"""GPU parity of the wav side through the C ABI: Vocoder_Taco1 (Modules/Taco2.py:234-260, CBHG :285-385) against the
golden produced by the reference's own source and against the fp64 oracle on other shapes and hyper-parameters; Griffin-Lim /
Audio.inv_spectrogram (Audio.py:23-27, 57-68) against the reference-source golden and the oracle, ragged batches included.
Tolerances: 1e-4 (fp32 handle) / 1e-2 (tensor-core handle) absolute on the spectrogram, as for the other layers; the waveform
is compared relative to its peak (Griffin-Lim iterates 60 STFT round trips in fp32)."""
import os

import numpy as np
import pytest
import torch

from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.weights import init_vocoder_weights, init_weights
from oracle import audio_port as A
from oracle import reference_port as O
from tests.util import BF16_TOL, FP32_TOL, max_abs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vocoder")


def _engine(cfg, seed=1357):
    from gst_tacotron_b200.runtime import Engine
    W = dict(init_weights(cfg, seed=1))
    WV = init_vocoder_weights(cfg, seed=seed)
    W.update(WV)
    return Engine(cfg, W), WV


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def eng(request):
    cfg = load_config(precision=request.param)
    e, WV = _engine(cfg)
    yield cfg, WV, e, (FP32_TOL if request.param == "fp32" else BF16_TOL)
    e.close()


def test_vocoder_matches_reference_golden(eng):
    cfg, WV, e, tol = eng
    z = np.load(os.path.join(GOLD, "vocoder.npz"))
    assert int(z["vocoder_seed"]) == 1357
    y = e.vocoder(z["mels"])
    assert y.shape == z["spectrogram"].shape
    assert max_abs(y, z["spectrogram"]) < tol


@pytest.mark.parametrize("B,T", [(1, 1), (3, 50), (2, 131), (5, 7)])
def test_vocoder_matches_oracle(eng, B, T):
    cfg, WV, e, tol = eng
    mels = (np.random.default_rng(B * 100 + T).standard_normal((B, T, cfg.mel_dim)) * 1.5).astype(np.float32)
    ref = O.vocoder(WV, cfg, mels)
    y = e.vocoder(torch.as_tensor(mels, device="cuda:0"))
    assert y.is_cuda and tuple(y.shape) == ref.shape
    assert max_abs(y, ref) < tol
    y2 = e.vocoder(mels)                       # host buffers in -> host buffers out, same numbers
    assert isinstance(y2, np.ndarray) and np.array_equal(y2, y.cpu().numpy())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vocoder_other_hyper_parameters(precision):
    """no Dense after the projections (last filter == Mel_Dim), no Dense before the Highwaynet (size == Mel_Dim), odd and even
    kernel sizes, a pool of 3, other widths"""
    cfg = load_config(precision=precision, voc_bank_count=5, voc_bank_filters=64, voc_pool_size=3, voc_proj_filters=[96, 80],
                      voc_proj_kernel=[4, 3], voc_highway_count=2, voc_highway_size=80, voc_rnn_size=128, spectrogram_dim=257)
    e, WV = _engine(cfg, seed=5)
    try:
        assert not any("Conv1D_Projection/dense" in k or "Highwaynet/dense" in k for k in WV)
        mels = (np.random.default_rng(3).standard_normal((3, 29, cfg.mel_dim)) * 1.5).astype(np.float32)
        ref = O.vocoder(WV, cfg, mels)
        assert max_abs(e.vocoder(mels), ref) < (FP32_TOL if precision == "fp32" else BF16_TOL)
    finally:
        e.close()


def test_vocoder_rejects_what_the_reference_cannot_run(eng):
    cfg, WV, e, tol = eng
    from gst_tacotron_b200.runtime import Engine
    cfg2 = load_config(precision=cfg.precision, voc_pool_strides=2)
    e2 = Engine(cfg2, init_weights(cfg2, seed=1))
    try:
        with pytest.raises(ValueError, match="strides"):
            e2.vocoder(np.zeros((1, 8, cfg.mel_dim), np.float32))
    finally:
        e2.close()
    from gst_tacotron_b200 import _lib
    e3 = Engine(cfg, init_weights(cfg, seed=1))       # no vocoder variables loaded
    try:
        with pytest.raises(_lib.GstkError, match="has not been loaded"):
            e3.vocoder(np.zeros((1, 8, cfg.mel_dim), np.float32))
    finally:
        e3.close()


# ---------------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


def test_inv_spectrogram_matches_reference_golden(eng):
    cfg, WV, e, tol = eng
    z = np.load(os.path.join(GOLD, "audio.npz"))
    for tag in "abc":
        spec, u, wav = z[tag + "_spec"], z[tag + "_uniform"], z[tag + "_wav"]
        mav = float(z[tag + "_max_abs"])
        F = spec.shape[0]
        got = e.griffin_lim(np.ascontiguousarray(spec.T[None]), iters=int(z[tag + "_iters"]), rng="external",
                            init_uniform=np.ascontiguousarray(u.T[None].astype(np.float32)), max_abs_value=None if mav < 0 else mav,
                            hop_length=(F - 1) // 2, win_length=(F - 1) * 2)
        assert got.shape == (1, wav.shape[0])
        assert _rel(got[0], wav) < 2e-3, tag


def test_single_pass_and_one_iteration_are_tight(eng):
    cfg, WV, e, tol = eng
    rng = np.random.default_rng(8)
    F, T = 513, 11
    spec = rng.uniform(-4, 4, (1, T, F)).astype(np.float32)
    u = rng.random((1, T, F)).astype(np.float32)
    for iters in (0, 1):
        ref = A.inv_spectrogram(spec[0].T, F, 256, 1024, 16000, max_abs_value=4, griffin_lim_iters=iters, init_uniform=u[0].T.astype(np.float64))
        got = e.griffin_lim(spec, iters=iters, rng="external", init_uniform=u, max_abs_value=4)
        assert _rel(got[0], ref) < 2e-4, iters


def test_ragged_batch_equals_one_utterance_at_a_time(eng):
    """Export_Inference cuts every utterance at its stop index before inv_spectrogram (Model.py:380,413): lengths[b] frames each"""
    cfg, WV, e, tol = eng
    rng = np.random.default_rng(12)
    F, T, B = 513, 17, 4
    lengths = np.array([17, 6, 2, 1], np.int32)
    spec = rng.uniform(-4, 4, (B, T, F)).astype(np.float32)
    u = rng.random((B, T, F)).astype(np.float32)
    got = e.griffin_lim(torch.as_tensor(spec, device="cuda:0"), lengths=lengths, iters=4, rng="external",
                        init_uniform=torch.as_tensor(u, device="cuda:0"), max_abs_value=4).cpu().numpy()
    assert got.shape == (B, 256 * (T - 1))
    for b in range(B):
        n = int(lengths[b])
        L = 256 * (n - 1)
        assert not got[b, L:].any()
        if n < 2:
            continue     # librosa cannot frame an empty signal; the library returns silence
        ref = A.inv_spectrogram(spec[b, :n].T, F, 256, 1024, 16000, max_abs_value=4, griffin_lim_iters=4, init_uniform=u[b, :n].T.astype(np.float64))
        assert _rel(got[b, :L], ref) < 1e-3, b
        alone = e.griffin_lim(spec[b:b + 1, :n], iters=4, rng="external", init_uniform=u[b:b + 1, :n], max_abs_value=4)
        assert np.array_equal(alone[0], got[b, :L])          # bit-identical to the utterance on its own


def test_philox_phases(eng):
    cfg, WV, e, tol = eng
    rng = np.random.default_rng(2)
    F, T, B = 129, 9, 3
    spec = rng.uniform(0, 1, (B, T, F)).astype(np.float32)
    kw = dict(iters=2, hop_length=64, win_length=256)
    a = e.griffin_lim(spec, rng="philox", seed=77, **kw)
    b = e.griffin_lim(spec, rng="philox", seed=77, **kw)
    c = e.griffin_lim(spec, rng="philox", seed=78, **kw)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    u = O.philox_uniform(77, O.STREAM_GRIFFIN_LIM, T, B, F)
    d = e.griffin_lim(spec, rng="external", init_uniform=u, **kw)
    assert np.array_equal(a, d)
    part = e.griffin_lim(spec[1:], rng="philox", seed=77, row_offset=1, **kw)     # sharded: rows keyed by their global index
    assert np.array_equal(part, a[1:])


def test_griffin_lim_argument_checks(eng):
    cfg, WV, e, tol = eng
    spec = np.zeros((1, 4, 513), np.float32)
    with pytest.raises(ValueError, match="init_uniform"):
        e.griffin_lim(spec, rng="external")
    with pytest.raises(ValueError, match="power of two"):
        e.griffin_lim(np.zeros((1, 4, 500), np.float32))
    with pytest.raises(ValueError, match="win_length"):
        e.griffin_lim(spec, win_length=800)
    with pytest.raises(ValueError, match="hop_length"):
        e.griffin_lim(spec, hop_length=300)


def test_reference_call_signatures(eng):
    """Modules.Taco2.Vocoder_Taco1()(inputs, training) and Audio.inv_spectrogram(spectrogram [num_freq, frames], ...) as the
    reference calls them (Model.py:126-129, 412-420); Engine.inference(wav=True) end to end"""
    cfg, WV, e, tol = eng
    from gst_tacotron_b200 import Audio
    from gst_tacotron_b200.Modules.Taco2 import Vocoder_Taco1
    z = np.load(os.path.join(GOLD, "vocoder.npz"))
    voc = Vocoder_Taco1(engine=e)
    assert max_abs(voc(z["mels"], training=False), z["spectrogram"]) < tol
    with pytest.raises(NotImplementedError):
        voc(z["mels"], training=True)
    g = np.load(os.path.join(GOLD, "audio.npz"))
    wav = Audio.inv_spectrogram(spectrogram=g["a_spec"], num_freq=513, hop_length=256, win_length=1024, sample_rate=16000,
                                max_abs_value=4, griffin_lim_iters=int(g["a_iters"]), init_uniform=g["a_uniform"], engine=e)
    assert wav.shape == g["a_wav"].shape and _rel(wav, g["a_wav"]) < 2e-3
    many = Audio.inv_spectrograms(np.ascontiguousarray(np.stack([g["a_spec"].T, g["a_spec"].T])), [9, 5], 256, 1024, max_abs_value=4,
                                  griffin_lim_iters=2, engine=e)
    assert [len(w) for w in many] == [256 * 8, 256 * 4]


def test_inference_to_waveform():
    from gst_tacotron_b200.runtime import Engine
    from gst_tacotron_b200.weights import init_encoder_weights, init_postnet_weights
    cfg = load_config(precision="bf16")
    W = dict(init_weights(cfg, seed=1))
    for part in (init_postnet_weights(cfg), init_encoder_weights(cfg), init_vocoder_weights(cfg)):
        W.update(part)
    e = Engine(cfg, W)
    try:
        rng = np.random.default_rng(0)
        tokens = rng.integers(0, cfg.vocab_size, (2, 11)).astype(np.int32)
        ref_mels = rng.standard_normal((2, 40, cfg.mel_dim)).astype(np.float32)
        out = e.inference(tokens, ref_mels, np.array([40, 33], np.int32), steps=12, seed=3, wav=True)
        assert out["spectrogram"].shape == (2, 12, cfg.spectrogram_dim) and out["wav"].shape == (2, cfg.frame_shift * 11)
        assert np.isfinite(out["wav"]).all() and np.isfinite(out["spectrogram"]).all()
        for b in range(2):
            stop = out["stop"][b]
            idx = int(np.argmax(stop < 0))                       # Model.py:380
            assert int(out["stop_index"][b]) == idx
            L = cfg.frame_shift * (max(1, idx) * cfg.step_reduction - 1)
            assert not out["wav"][b, max(L, 0):].any()
    finally:
        e.close()


def test_highway_stack_both_kernels_in_tensor_core_mode(monkeypatch):
    """tensor-core handle: mma.sync stack (default) and the FFMA stack on fp16 matrices (GSTK_VOC_HIGHWAY=ffma) both within 1e-2"""
    cfg = load_config(precision="bf16")
    e, WV = _engine(cfg)
    try:
        mels = (np.random.default_rng(17).standard_normal((2, 77, cfg.mel_dim)) * 1.5).astype(np.float32)
        ref = O.vocoder(WV, cfg, mels)
        a = e.vocoder(mels)
        monkeypatch.setenv("GSTK_VOC_HIGHWAY", "ffma")
        b = e.vocoder(mels)
        assert max_abs(a, ref) < BF16_TOL and max_abs(b, ref) < BF16_TOL
        assert not np.array_equal(a, b)
    finally:
        e.close()


def test_wav_side_at_the_benchmark_batch():
    """256 utterances x 1000 frames (BASELINE configs[2]'s decode output) through Vocoder_Taco1 and Griffin-Lim in the tensor-core
    mode: utterances 0 and 255 against the fp64 oracle run on them alone (the oracle cannot run the whole batch in test time; the
    kernels treat utterances independently, which the comparison of both ends of the batch checks), everything finite, every
    utterance of the ragged Griffin-Lim silent beyond its own length."""
    cfg = load_config(precision="bf16")
    e, WV = _engine(cfg)
    try:
        B, T = 256, 1000
        g = torch.Generator(device="cuda").manual_seed(3)
        mels = torch.randn(B, T, cfg.mel_dim, device="cuda", generator=g) * 1.5
        spec = e.vocoder(mels)
        assert tuple(spec.shape) == (B, T, cfg.spectrogram_dim) and torch.isfinite(spec).all()
        for b in (0, 255):
            ref = O.vocoder(WV, cfg, mels[b:b + 1].cpu().numpy())
            assert max_abs(spec[b:b + 1], ref) < BF16_TOL, b
        lengths = torch.randint(2, T + 1, (B,), generator=torch.Generator().manual_seed(4)).to(torch.int32)
        lengths[0], lengths[255] = T, 37
        wav = e.griffin_lim(spec, lengths=lengths.cuda(), iters=3, rng="philox", seed=9, max_abs_value=cfg.max_abs_mel)
        assert tuple(wav.shape) == (B, cfg.frame_shift * (T - 1)) and torch.isfinite(wav).all()
        idx = torch.arange(wav.shape[1], device="cuda")[None, :]
        L = (cfg.frame_shift * (lengths.cuda() - 1))[:, None]
        assert not wav[idx >= L].any() and wav[idx < L].abs().max() > 0
        u = O.philox_uniform(9, O.STREAM_GRIFFIN_LIM, T, B, cfg.spectrogram_dim)
        for b in (0, 255):
            n = int(lengths[b])
            ref = A.inv_spectrogram(spec[b, :n].cpu().numpy().T, cfg.spectrogram_dim, cfg.frame_shift, cfg.frame_length, cfg.sample_rate,
                                    max_abs_value=cfg.max_abs_mel, griffin_lim_iters=3, init_uniform=u[b, :n].T.astype(np.float64))
            assert _rel(wav[b, :cfg.frame_shift * (n - 1)].cpu().numpy(), ref) < 1e-3, b
    finally:
        e.close()
