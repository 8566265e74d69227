"""Text Encoder (reference: Modules/Taco2.py:12-51) through the C ABI: against the golden vector produced
by the reference's own Sequential, against the CPU oracle on ragged sizes, and chained into the decoder.
fp32: 1e-4 absolute; tensor-core mode: 1e-2 (north_star tolerances)."""
import os

import numpy as np
import pytest
import torch

from tests.test_golden_cpu import GOLD, err
from tests.util import make_cfg, make_weights
from oracle import reference_port as O  # checker only

pytestmark = pytest.mark.gpu

TOLS = [("fp32", 1e-4), ("bf16", 1e-2)]


def _engine(cfg, precision, seed=2468, with_encoder=True):
    from gst_tacotron_b200.runtime import Engine
    from gst_tacotron_b200.weights import init_encoder_weights
    cfg.precision = precision
    W = dict(make_weights(cfg))
    WE = init_encoder_weights(cfg, seed=seed)
    if with_encoder:
        W.update(WE)
    return Engine(cfg, W), WE


@pytest.mark.parametrize("precision,tol", TOLS)
def test_encoder_against_reference_golden(precision, tol):
    g = np.load(os.path.join(GOLD, "encoder", "encoder.npz"))
    cfg = make_cfg()
    eng, _ = _engine(cfg, precision, seed=int(g["encoder_seed"]))
    try:
        got = eng.encoder(g["tokens"])
        assert got.shape == g["encodings"].shape
        assert err(got, g["encodings"]) < tol
    finally:
        eng.close()


@pytest.mark.parametrize("precision,tol", TOLS)
@pytest.mark.parametrize("B,Tv", [(1, 1), (1, 82), (3, 2), (5, 37), (7, 150)])
def test_encoder_matches_oracle_ragged_sizes(precision, tol, B, Tv):
    """key_time below the conv width, batch not a multiple of the 4-utterance LSTM group, BASELINE's T_v = 82 / 150."""
    cfg = make_cfg()
    eng, WE = _engine(cfg, precision)
    try:
        rng = np.random.default_rng(B * 1000 + Tv)
        tokens = rng.integers(0, cfg.vocab_size, size=(B, Tv)).astype(np.int32)
        want = O.encoder(WE, cfg, tokens)
        got_host = eng.encoder(tokens)
        got_dev = eng.encoder(torch.from_numpy(tokens).cuda())
        assert isinstance(got_host, np.ndarray) and got_dev.is_cuda
        assert err(got_host, want) < tol
        assert np.array_equal(got_host, got_dev.cpu().numpy())
    finally:
        eng.close()


def test_encoder_batch_independence_full_size():
    """BASELINE configs[2] size (256 utterances x 150 tokens), tensor-core mode: utterances do not see their neighbours."""
    cfg = make_cfg()
    eng, WE = _engine(cfg, "bf16")
    try:
        gen = torch.Generator(device="cuda").manual_seed(3)
        tokens = torch.randint(0, cfg.vocab_size, (256, 150), device="cuda", generator=gen, dtype=torch.int32)
        full = eng.encoder(tokens)
        torch.cuda.synchronize()
        ms = eng.last_kernel_ms()
        assert bool(torch.isfinite(full).all()) and float(full.abs().max()) <= 1.0
        for rows in ([0], [255], [101, 102, 103, 104, 105]):
            assert torch.equal(eng.encoder(tokens[rows].contiguous()), full[rows])
        want = O.encoder(WE, cfg, tokens[[0, 255]].cpu().numpy())
        assert err(full[[0, 255]].cpu().numpy(), want) < 1e-2
        print("encoder bf16 256x150: {:.2f} ms".format(ms))
    finally:
        eng.close()


@pytest.mark.parametrize("precision,tol", TOLS)
def test_encoder_dropin_feeds_decoder(precision, tol):
    """Encoder()(tokens) -> decode(enc_text=..., gst=...) equals the oracle's encoder -> GST concat -> decoder loop."""
    from gst_tacotron_b200.Modules.Taco2 import Encoder
    cfg = make_cfg()
    eng, WE = _engine(cfg, precision)
    try:
        B, Tv, T = 2, 17, 5
        rng = np.random.default_rng(11)
        tokens = rng.integers(0, cfg.vocab_size, size=(B, Tv)).astype(np.int32)
        gst = rng.uniform(-1, 1, size=(B, cfg.style_size)).astype(np.float32)
        _, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
        enc = Encoder(eng)(torch.from_numpy(tokens).cuda(), training=False)
        out = eng.decode(enc_text=enc, gst=torch.from_numpy(gst).cuda(), teacher_mels=torch.from_numpy(mels[:, :-1]).cuda(),
                         rng="external", keep0=k0, keep1=k1, noise=nz)
        enc_ref = O.encoder(WE, cfg, tokens)
        cat = np.concatenate([np.repeat(gst[:, None, :], Tv, axis=1), enc_ref], axis=-1)   # GST.py:121-124, GST channels first
        ref = O.decoder_loop(make_weights(cfg), cfg, cat, mels=mels, training=True, keep0=k0, keep1=k1, noise=nz)
        assert err(out["mel"].cpu().numpy(), ref["decodings"].numpy()) < tol
        assert err(out["alignment"].cpu().numpy(), ref["alignments"].numpy()) < tol
    finally:
        eng.close()


@pytest.mark.parametrize("which", ["stream", "persistent"])
def test_encoder_both_recurrent_kernels_match_oracle(which, monkeypatch):
    """Both recurrent kernels (persistent: the default when its grid fits the device; streaming: the fall-back for RNN sizes
    that do not) against the oracle (GSTK_ENC_BILSTM is read per call)."""
    monkeypatch.setenv("GSTK_ENC_BILSTM", which)
    cfg = make_cfg()
    eng, WE = _engine(cfg, "fp32")
    try:
        B, Tv = (3, 21) if which == "stream" else (259, 9)     # 259: two launches of the persistent kernel (256 + 3 utterances)
        tokens = np.random.default_rng(B).integers(0, cfg.vocab_size, size=(B, Tv)).astype(np.int32)
        assert err(eng.encoder(tokens), O.encoder(WE, cfg, tokens)) < 1e-4
    finally:
        eng.close()


def test_inference_chain_matches_oracle_chain():
    """Engine.inference = the reference's Inference model up to the vocoder (Model.py:108-125): tokens + reference mels ->
    Encoder -> Style_Token_Layer -> concat -> free-running Decoder -> Postnet, against the same chain of oracle functions."""
    from gst_tacotron_b200.runtime import Engine
    from gst_tacotron_b200.weights import init_encoder_weights, init_postnet_weights
    cfg = make_cfg()
    cfg.precision = "fp32"
    W, WE, WP = make_weights(cfg), init_encoder_weights(cfg), init_postnet_weights(cfg)
    eng = Engine(cfg, {**W, **WE, **WP})
    try:
        B, Tv, T = 2, 15, 6
        rng = np.random.default_rng(21)
        tokens = rng.integers(0, cfg.vocab_size, size=(B, Tv)).astype(np.int32)
        gm, gl = O.synth_gst_inputs(cfg, B, 96)
        _, _, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
        got = eng.inference(tokens, gm, gl, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
        enc = O.encoder(WE, cfg, tokens)
        style = O.style_token_layer(W, cfg, gm, gl).numpy()
        cat = np.concatenate([np.repeat(style[:, None, :], Tv, axis=1), enc], axis=-1)
        ref = O.decoder_loop(W, cfg, cat, training=False, steps=T, keep0=k0, keep1=k1, noise=nz)
        assert err(got["encodings"], enc) < 1e-4 and err(got["gst"], style) < 1e-4
        assert err(got["mel"], ref["decodings"].numpy()) < 1e-4
        assert err(got["alignment"], ref["alignments"].numpy()) < 1e-4
        assert err(got["post_mel"], O.postnet(WP, cfg, ref["decodings"].numpy())) < 1e-3
    finally:
        eng.close()


def test_encoder_errors():
    from gst_tacotron_b200._lib import GstkError
    cfg = make_cfg()
    eng, _ = _engine(cfg, "fp32", with_encoder=False)
    try:
        assert not eng.has_encoder
        with pytest.raises(GstkError) as ei:
            eng.encoder(np.zeros((1, 4), np.int32))
        assert ei.value.code == 4
        with pytest.raises(ValueError):
            eng.encoder(np.zeros((4,), np.int32))
    finally:
        eng.close()


@pytest.mark.parametrize("precision,tol", TOLS)
def test_non_default_encoder_and_postnet_shapes(precision, tol):
    """Hyper-parameters away from the defaults: even conv kernels (TF 'same' pads 1 before / 2 after for k = 4), channel counts
    that are not multiples of 64 (overlapping-row tensor maps with a zero-filled K tail in the tensor-core mode), an RNN size
    that takes the FFMA persistent recurrence in both modes."""
    from gst_tacotron_b200.hparams import config_from_hp
    from gst_tacotron_b200.runtime import Engine
    from gst_tacotron_b200.weights import init_encoder_weights, init_postnet_weights
    cfg = config_from_hp({"Tacotron2": {
        "Encoder": {"Embedding": {"Size": 128}, "Conv": {"Filters": [64, 96], "Kernel_Size": [3, 4], "Strides": [1, 1]},
                    "RNN": {"Size": 64}},
        "Decoder": {"Conv": {"Filters": [32, 48], "Kernel_Size": [4, 5], "Strides": [1, 1]}}}})
    cfg.precision = precision
    WE, WP = init_encoder_weights(cfg), init_postnet_weights(cfg)
    eng = Engine(cfg, {**make_weights(cfg), **WE, **WP})
    try:
        rng = np.random.default_rng(8)
        tokens = rng.integers(0, cfg.vocab_size, size=(5, 23)).astype(np.int32)
        got = eng.encoder(tokens)
        assert got.shape == (5, 23, 128)
        assert err(got, O.encoder(WE, cfg, tokens)) < tol
        dec = rng.uniform(-1, 1, size=(3, 41, cfg.mel_dim)).astype(np.float32)
        assert err(eng.postnet(dec), O.postnet(WP, cfg, dec)) < tol
    finally:
        eng.close()
