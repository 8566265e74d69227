"""The CUDA path (through the C ABI) against the golden vectors generated from the reference's own sources
(tests/golden/*.npz, see oracle/make_golden.py).  fp32: 1e-4 absolute; bf16: 1e-2 (north_star tolerances)."""
import numpy as np
import pytest
import torch

from tests.test_golden_cpu import NAMES, err, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from gst_tacotron_b200.runtime import Engine
    cache = {}

    def get(name, precision):
        if (name, precision) not in cache:
            g, cfg, W = load(name)
            cfg.precision = precision
            cache[(name, precision)] = (g, cfg, Engine(cfg, W))
        return cache[(name, precision)]
    yield get
    for _, _, e in cache.values():
        e.close()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("name", NAMES)
def test_decoder_against_reference_goldens(engines, name, precision, tol):
    g, cfg, eng = engines(name, precision)
    r = cfg.step_reduction
    out = eng.decode(encodings=g["tf_enc"], teacher_mels=g["tf_mels"][:, 0:-1:r], rng="external", keep0=g["tf_keep0"],
                     keep1=g["tf_keep1"], noise=g["tf_noise"])
    assert err(out["mel"], g["tf_decodings"]) < tol
    assert err(out["stop"], g["tf_stops"]) < tol
    assert err(out["alignment"], g["tf_alignments"]) < tol
    fr = eng.decode(encodings=g["tf_enc"], steps=cfg.max_step // r, rng="external", keep0=g["fr_keep0"], keep1=g["fr_keep1"],
                    noise=g["fr_noise"])
    assert fr["mel"].shape == g["fr_decodings"].shape
    assert err(fr["mel"], g["fr_decodings"]) < tol
    assert err(fr["alignment"], g["fr_alignments"]) < tol
    clear = np.abs(g["fr_stops"]) > tol  # stop-frame identity (Model.py:380) wherever the sign is decidable
    assert np.array_equal((np.asarray(fr["stop"]) < 0)[clear], (g["fr_stops"] < 0)[clear])


@pytest.mark.parametrize("name", NAMES)
def test_decoder_step_layer_against_reference_goldens(engines, name):
    """Decoder_Step drop-in (Taco2.py:96-120 signature) on the reference's single-step golden."""
    from gst_tacotron_b200.Modules.Taco2 import Decoder_Step
    g, cfg, eng = engines(name, "fp32")
    s = g["step_states_in"]
    ds = Decoder_Step(eng)
    mel, stop, al, st = ds([g["step_enc"], g["step_mel_in"], g["step_prev_alignment"], ((s[0], s[1]), (s[2], s[3]))],
                           training=False, rng="external", keep0=g["step_keep0"][0], keep1=g["step_keep1"][0],
                           noise=g["step_noise"][0])
    assert err(mel.cpu(), g["step_mel"]) < 1e-4 and err(stop.cpu(), g["step_stop"]) < 1e-4
    assert err(al.cpu(), g["step_alignment"]) < 1e-4
    got = np.stack([st[0][0].cpu().numpy(), st[0][1].cpu().numpy(), st[1][0].cpu().numpy(), st[1][1].cpu().numpy()])
    assert err(got, g["step_states"]) < 1e-4


@pytest.mark.parametrize("name", NAMES)
def test_gst_against_reference_goldens(engines, name):
    from gst_tacotron_b200.Modules.GST import GST_Concated_Encoder, Reference_Encoder, Style_Token_Layer
    g, cfg, eng = engines(name, "fp32")
    style = Style_Token_Layer(eng)([g["gst_mels"], g["gst_lengths"]])
    ref = Reference_Encoder(eng)([np.ascontiguousarray(g["gst_mels"][:, 1:]), g["gst_lengths"]])
    assert err(ref, g["gst_ref"]) < 1e-4
    assert err(style, g["gst_style"]) < 1e-4
    cat = GST_Concated_Encoder(eng)([np.ascontiguousarray(g["tf_enc"][:, :, cfg.style_size:]),
                                     np.ascontiguousarray(g["tf_enc"][:, 0, :cfg.style_size])])
    assert err(cat, g["cat_out"]) == 0.0
    out, dist = eng.mha(g["mha_q"], g["mha_v"], g["mha_Query_kernel"], g["mha_Query_bias"], g["mha_Value_kernel"],
                        g["mha_Value_bias"], g["mha_gamma"], g["mha_beta"], 8)
    assert err(out, g["mha_out"]) < 1e-4 and err(dist, g["mha_dist"]) < 1e-4
