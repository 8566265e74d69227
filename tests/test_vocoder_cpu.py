"""CPU checks of the oracle for the wav side: Vocoder_Taco1 (Modules/Taco2.py:234-260, CBHG :285-385) and
Audio.inv_spectrogram / Griffin-Lim (Audio.py:23-27, 57-68) against the goldens produced by running the reference's own sources
(oracle/make_golden.py --vocoder / --audio), and the restated librosa stft / istft against torch's."""
import os

import numpy as np
import torch

from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.weights import init_vocoder_weights, vocoder_spec
from oracle import audio_port as A
from oracle import reference_port as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vocoder")


def test_oracle_vocoder_matches_reference_golden():
    cfg = load_config()
    z = np.load(os.path.join(GOLD, "vocoder.npz"))
    W = init_vocoder_weights(cfg, seed=int(z["vocoder_seed"]))
    assert list(W) == list(vocoder_spec(cfg))
    y = O.vocoder(W, cfg, z["mels"])
    assert y.shape == z["spectrogram"].shape == (2, 21, cfg.spectrogram_dim)
    assert np.abs(y - z["spectrogram"]).max() < 1e-9
    assert np.abs(z["spectrogram"]).max() > 0.5     # the golden is not a field of zeros


def test_max_pool_same_is_max_with_the_next_frame():
    x = torch.randn(2, 7, 3, dtype=torch.float64)
    y = O.max_pool1d_same(x, 2, 1)
    ref = torch.maximum(x, torch.cat([x[:, 1:], torch.full((2, 1, 3), float("-inf"), dtype=torch.float64)], 1))
    assert torch.equal(y, ref)


def test_stft_istft_restatement_matches_torch():
    rng = np.random.default_rng(0)
    for n_fft, hop, frames in ((1024, 256, 9), (128, 32, 5), (1024, 256, 4)):
        y = rng.standard_normal(hop * (frames - 1))
        w = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
        D = A.stft(y, n_fft, hop)
        Dt = torch.stft(torch.as_tensor(y), n_fft, hop, n_fft, w, center=True, pad_mode="reflect", return_complex=True).numpy()
        assert D.shape == Dt.shape == (n_fft // 2 + 1, frames)
        assert np.abs(D - Dt).max() < 1e-9
        X = rng.standard_normal(D.shape) + 1j * rng.standard_normal(D.shape)
        yi = A.istft(X, hop, n_fft)
        yt = torch.istft(torch.as_tensor(X), n_fft, hop, n_fft, w, center=True).numpy()
        assert yi.shape == yt.shape == (hop * (frames - 1),)
        assert np.abs(yi - yt).max() < 1e-9


def test_inv_spectrogram_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "audio.npz"))
    for tag in "abc":
        spec, u, wav = z[tag + "_spec"], z[tag + "_uniform"], z[tag + "_wav"]
        mav = float(z[tag + "_max_abs"])
        F = spec.shape[0]
        got = A.inv_spectrogram(spec, F, (F - 1) // 2, (F - 1) * 2, 16000, max_abs_value=None if mav < 0 else mav,
                                griffin_lim_iters=int(z[tag + "_iters"]), init_uniform=u)
        assert got.shape == wav.shape
        scale = np.abs(wav).max()
        assert scale > 0
        assert np.abs(got - wav).max() < 1e-7 * max(scale, 1.0), tag


def test_vocoder_pack_by_layer_order():
    """weights.vocoder_pack: the variables handed over layer by layer (as a dump of the reference's Keras objects would give them)
    reproduce the canonical pack; a wrong count or a Dense that the configuration does not have is an error."""
    import pytest
    from gst_tacotron_b200.weights import VOC, vocoder_pack
    cfg = load_config()
    W = init_vocoder_weights(cfg)
    bn = lambda base: tuple(W[base + leaf] for leaf in ("gamma", "beta", "moving_mean", "moving_variance"))
    bank = [(W[VOC + "/CBHG/ConvBank_%d/conv1d/kernel" % i], bn(VOC + "/CBHG/ConvBank_%d/batch_normalization/" % i)) for i in range(cfg.voc_bank_count)]
    proj = [(W[VOC + "/CBHG/Conv1D_Projection/conv1d_%d/kernel" % i], bn(VOC + "/CBHG/Conv1D_Projection/batch_normalization_%d/" % i))
            for i in range(len(cfg.voc_proj_filters))]
    kb = lambda base: (W[base + "kernel"], W[base + "bias"])
    hws = [(kb(VOC + "/CBHG/Highwaynet/highwaynet_%d/Dense_Relu/" % i), kb(VOC + "/CBHG/Highwaynet/highwaynet_%d/Dense_Sigmoid/" % i))
           for i in range(cfg.voc_highway_count)]
    cell = lambda d: tuple(W[VOC + "/CBHG/RNN/%s/lstm_cell/%s" % (d, leaf)] for leaf in ("kernel", "recurrent_kernel", "bias"))
    got = vocoder_pack(cfg, bank, proj, kb(VOC + "/CBHG/Conv1D_Projection/dense/"), kb(VOC + "/CBHG/Highwaynet/dense/"), hws,
                       cell("forward_lstm"), cell("backward_lstm"), kb(VOC + "/Dense/"))
    assert list(got) == list(vocoder_spec(cfg)) and all(np.array_equal(got[k], W[k]) for k in W)
    with pytest.raises(ValueError):
        vocoder_pack(cfg, bank[:-1], proj, kb(VOC + "/CBHG/Conv1D_Projection/dense/"), kb(VOC + "/CBHG/Highwaynet/dense/"), hws,
                     cell("forward_lstm"), cell("backward_lstm"), kb(VOC + "/Dense/"))
    with pytest.raises(ValueError):
        vocoder_pack(cfg, bank, proj, None, kb(VOC + "/CBHG/Highwaynet/dense/"), hws, cell("forward_lstm"), cell("backward_lstm"), kb(VOC + "/Dense/"))
