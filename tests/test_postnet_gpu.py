"""Postnet (reference: Modules/Taco2.py:130-149 + the residual add of :230) through the C ABI:
against the golden vectors produced by the reference's own Sequential, against the CPU oracle on ragged sizes, and through
size-independent properties at BASELINE's full size.  fp32: 1e-4 absolute; bf16: 1e-2 (north_star tolerances)."""
import numpy as np
import pytest
import torch

from tests.test_golden_cpu import NAMES, err, load
from tests.util import make_cfg, make_weights
from oracle import reference_port as O  # checker only

pytestmark = pytest.mark.gpu

TOLS = [("fp32", 1e-4), ("bf16", 1e-2)]


def _engine(cfg, precision, postnet_seed=4321, with_postnet=True, W=None):
    from gst_tacotron_b200.runtime import Engine
    from gst_tacotron_b200.weights import init_postnet_weights
    cfg.precision = precision
    W = dict(make_weights(cfg) if W is None else W)
    WP = init_postnet_weights(cfg, seed=postnet_seed)
    if with_postnet:
        W.update(WP)
    return Engine(cfg, W), WP


@pytest.mark.parametrize("precision,tol", TOLS)
@pytest.mark.parametrize("name", NAMES)
def test_postnet_against_reference_goldens(name, precision, tol):
    g, cfg, W = load(name)
    eng, _ = _engine(cfg, precision, postnet_seed=int(g["postnet_seed"]), W=W)
    try:
        got = eng.postnet(g["fr_decodings"].astype(np.float32))
        assert got.shape == g["fr_post_decodings"].shape
        assert err(got, g["fr_post_decodings"]) < tol
    finally:
        eng.close()


@pytest.mark.parametrize("precision,tol", TOLS)
@pytest.mark.parametrize("B,T", [(1, 1), (2, 3), (3, 333), (5, 128), (2, 1000)])
def test_postnet_matches_oracle_ragged_sizes(precision, tol, B, T):
    """Frame counts below the kernel width, not a multiple of the 128-row CTA tile, utterance boundaries inside a tile."""
    cfg = make_cfg()
    eng, WP = _engine(cfg, precision)
    try:
        rng = np.random.default_rng(B * 1000 + T)
        dec = rng.uniform(-4, 4, size=(B, T, cfg.mel_dim)).astype(np.float32)
        want = O.postnet(WP, cfg, dec)
        got_host = eng.postnet(dec)                                    # host pointers (staged by the library)
        got_dev = eng.postnet(torch.from_numpy(dec).cuda())            # device pointers (borrowed)
        assert isinstance(got_host, np.ndarray) and got_dev.is_cuda
        assert err(got_host, want) < tol
        assert np.array_equal(got_host, got_dev.cpu().numpy())
    finally:
        eng.close()


def test_postnet_full_size_properties():
    """BASELINE configs[2] size (256 utterances x 1000 frames), bf16 tensor-core kernels: every utterance is independent of
    its neighbours in the flat padded matrix (bitwise), rows agree with the oracle, output is finite."""
    cfg = make_cfg()
    eng, WP = _engine(cfg, "bf16")
    try:
        B, T = 256, 1000
        gen = torch.Generator(device="cuda").manual_seed(5)
        dec = (torch.rand(B, T, cfg.mel_dim, device="cuda", generator=gen) * 8 - 4)
        full = eng.postnet(dec)
        torch.cuda.synchronize()
        ms = eng.last_kernel_ms()
        assert bool(torch.isfinite(full).all())
        for rows in ([0], [255], [17, 18, 19]):
            alone = eng.postnet(dec[rows].contiguous())
            assert torch.equal(alone, full[rows])
        want = O.postnet(WP, cfg, dec[[0, 255]].cpu().numpy())
        assert err(full[[0, 255]].cpu().numpy(), want) < 1e-2
        flops = 2.0 * B * T * sum(k * ci * co for (co, k, _s, _t), ci in
                                  zip(cfg.postnet_layers, [cfg.mel_dim] + [l[0] for l in cfg.postnet_layers[:-1]]))
        print("postnet bf16 256x1000: {:.2f} ms, {:.1f} TFLOP/s".format(ms, flops / ms * 1e-9))
    finally:
        eng.close()


def test_postnet_fp32_independent_of_batch_position():
    cfg = make_cfg()
    eng, _ = _engine(cfg, "fp32")
    try:
        gen = torch.Generator(device="cuda").manual_seed(6)
        dec = (torch.rand(8, 300, cfg.mel_dim, device="cuda", generator=gen) * 8 - 4)
        full = eng.postnet(dec)
        assert torch.equal(eng.postnet(dec[5:6].contiguous()), full[5:6])
    finally:
        eng.close()


@pytest.mark.parametrize("precision,tol", TOLS)
def test_decoder_dropin_returns_post_decodings(precision, tol):
    """Decoder.call's 4-tuple (Taco2.py:232): post_decodings = Postnet(decodings) + decodings of ITS OWN decodings."""
    from gst_tacotron_b200.Modules.Taco2 import Decoder, Postnet
    cfg = make_cfg()
    eng, WP = _engine(cfg, precision)
    try:
        B, Tv, T = 3, 21, 9
        enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
        ins = [torch.from_numpy(np.asarray(enc, np.float32)).cuda(), torch.from_numpy(np.asarray(mels, np.float32)).cuda()]
        dec, post, stops, align = Decoder(eng)(ins, training=False, max_steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
        assert post is not None and post.shape == dec.shape
        # training=True: the reference's Postnet then runs batch-statistics BatchNorm + Dropout (Taco2.py:144-149), which is
        # not built here - the drop-in returns None for it instead of the inference-form tensor
        assert Decoder(eng)(ins, training=True, rng="external", keep0=k0, keep1=k1, noise=nz)[1] is None
        assert err(post.cpu().numpy(), O.postnet(WP, cfg, dec.cpu().numpy())) < tol
        assert torch.equal(Postnet(eng)(dec), post)
    finally:
        eng.close()


def test_postnet_errors():
    from gst_tacotron_b200._lib import GstkError
    cfg = make_cfg()
    eng, _ = _engine(cfg, "fp32", with_postnet=False)
    try:
        assert not eng.has_postnet
        with pytest.raises(GstkError) as ei:                           # decode-only pack: variables not loaded
            eng.postnet(np.zeros((1, 4, cfg.mel_dim), np.float32))
        assert ei.value.code == 4
        with pytest.raises(ValueError):
            eng.postnet(np.zeros((1, 4, cfg.mel_dim + 1), np.float32))
    finally:
        eng.close()
    cfg2 = make_cfg()
    cfg2.postnet_strides = [1, 2, 1, 1]
    eng2, _ = _engine(cfg2, "fp32")
    try:
        with pytest.raises(ValueError):
            eng2.postnet(np.zeros((1, 4, cfg2.mel_dim), np.float32))
    finally:
        eng2.close()
