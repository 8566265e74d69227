"""CPU-only checks of the oracle against facts that follow directly from the reference code
 and of the shared Philox stream."""
import numpy as np
import torch

from oracle import reference_port as O
from tests.util import make_cfg, make_weights


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    r = O.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = O.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x) for x in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = O.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_streams_statistics():
    cfg = make_cfg()
    k0, k1, nz = O.philox_randomness(cfg, seed=42, T=20, B=8, Tv=150)
    assert k0.shape == (20, 8, 256) and nz.shape == (20, 8, 150)
    assert abs(k0.mean() - 0.5) < 0.02 and abs(k1.mean() - 0.5) < 0.02
    assert abs(nz.mean()) < 0.03 and abs(nz.std() - 1.0) < 0.03
    # offsets address the same stream
    k0b, _, nzb = O.philox_randomness(cfg, seed=42, T=5, B=3, Tv=150, t0=7, b0=2)
    assert np.array_equal(k0b, k0[7:12, 2:5]) and np.array_equal(nzb, nz[7:12, 2:5])


def test_initial_alignment_and_sma_mass():
    cfg = make_cfg("SMA")
    W = make_weights(cfg)
    B, Tv, T = 2, 30, 12
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    out = O.decoder_loop(W, cfg, enc, mels=mels, training=True, keep0=k0, keep1=k1, noise=nz)
    al = out["alignments"].numpy()
    assert out["decodings"].shape == (B, T, 80) and out["stops"].shape == (B, T) and al.shape == (B, T, Tv)
    assert np.allclose(al.sum(-1), 1.0, atol=1e-12)  # T < Tv: nothing can leak off the right edge
    assert np.all(al[:, 0, 2:] == 0)                # one-hot start, at most one move per step
    a0 = O.initial_alignment(cfg, B, Tv, torch.float64).numpy()
    assert np.array_equal(a0[:, 0], np.ones(B)) and a0[:, 1:].sum() == 0


def test_bma_matches_direct_cumprod():
    """safe_cumprod in log space == plain exclusive cumprod when nothing underflows (Steps.py:183-198)."""
    rng = np.random.default_rng(0)
    score = torch.as_tensor(rng.standard_normal((3, 17)))
    prev = torch.softmax(torch.as_tensor(rng.standard_normal((3, 17))), -1)
    got = O.bma_probability(score, prev, None, 0.0).numpy()
    p = 1 / (1 + np.exp(-score.numpy()))
    cp = np.cumprod(np.concatenate([np.ones((3, 1)), (1 - p)[:, :-1]], 1), axis=1)
    want = p * cp * np.cumsum(prev.numpy() / np.clip(cp, 1e-10, 1.0), axis=1)
    assert np.allclose(got, want, atol=1e-12)


def test_conv_same_padding_matches_tf_rule():
    # even input: pad (0,1); odd input: pad (1,1)  -> out = ceil(in/2)
    for n in (80, 5, 3, 188, 47):
        x = torch.arange(n * n, dtype=torch.float64).reshape(1, n, n, 1)
        k = torch.zeros(3, 3, 1, 1, dtype=torch.float64)
        k[0, 0, 0, 0] = 1.0  # picks the top-left tap => reveals pad_before
        y = O.conv2d_same_nhwc(x, k, 2)
        assert y.shape[1] == -(-n // 2)
        first = y[0, 0, 0, 0].item()
        assert first == (0.0 if n % 2 else x[0, 0, 0, 0].item())
        if n % 2 == 0:
            assert y[0, 1, 1, 0].item() == x[0, 2, 2, 0].item()


def test_style_token_layer_shapes_and_layernorm():
    cfg = make_cfg()
    W = make_weights(cfg)
    mels, lens = O.synth_gst_inputs(cfg, 3, 130)
    out, ref, att = O.style_token_layer(W, cfg, mels, lens, return_parts=True)
    assert out.shape == (3, 128) and ref.shape == (3, 128) and att.shape == (3, 16)
    assert np.allclose(att.sum(-1).numpy(), 1.0)
    assert np.allclose(out.mean(-1).numpy(), 0.0, atol=1e-9)
    assert np.allclose(out.var(-1, unbiased=False).numpy(), 1.0, atol=1e-4)


def test_gst_concat_order():
    e = torch.zeros(2, 5, 4)
    g = torch.ones(2, 3)
    c = O.gst_concat(e, g)
    assert c.shape == (2, 5, 7) and c[..., :3].min() == 1 and c[..., 3:].max() == 0


def test_free_running_feeds_last_frame_back():
    cfg = make_cfg("SMA", step_reduction=2)
    W = make_weights(cfg)
    enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, 1, 20, 3, teacher=False)
    full = O.decoder_loop(W, cfg, enc, steps=3, keep0=k0, keep1=k1, noise=nz)
    # teacher-forcing with the produced frames (last of each r) reproduces the free run
    prod = full["decodings"].numpy()
    teach = np.zeros((1, 3 * 2 + 1, 80), np.float32)
    teach[0, 2] = prod[0, 1]
    teach[0, 4] = prod[0, 3]
    tf = O.decoder_loop(W, cfg, enc, mels=teach, training=True, keep0=k0, keep1=k1, noise=nz)
    assert np.allclose(tf["decodings"].numpy(), prod, atol=1e-6)
