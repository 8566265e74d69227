"""The CPU oracle against golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCES
(oracle/make_golden.py: /root/reference/Modules/*.py on the torch-backed TensorFlow shim).  This pins the
oracle for everything that lives in the reference's Python code; tolerance is float64 round-off."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from gst_tacotron_b200.hparams import DEFAULT_HP, config_from_hp
from gst_tacotron_b200.weights import init_weights
from oracle import reference_port as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
TOL = 1e-9


def load(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    over = json.load(open(os.path.join(GOLD, name + ".hp.json")))
    hp = json.loads(json.dumps(DEFAULT_HP))
    for k, v in over.items():
        d = hp
        ks = k.split(".")
        for kk in ks[:-1]:
            d = d[kk]
        d[ks[-1]] = v
    cfg = config_from_hp(hp)
    W = init_weights(cfg, seed=int(g["weights_seed"]), bias_scale=float(g["bias_scale"]))
    return g, cfg, W


def err(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def test_golden_files_present():
    assert set(NAMES) >= {"sma_r1", "bma_r1", "sma_r2", "gst10"}


@pytest.mark.parametrize("name", NAMES)
def test_decoder_step_matches_reference_sources(name):
    g, cfg, W = load(name)
    Wt = O.to_torch(W)
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    values = O.value_projection(Wt, t(g["step_enc"]))
    s = g["step_states_in"]
    mel, stop, al, st, _ = O.decoder_step(Wt, cfg, values, t(g["step_mel_in"]), t(g["step_prev_alignment"]),
                                          ((t(s[0]), t(s[1])), (t(s[2]), t(s[3]))), t(g["step_keep0"][0]), t(g["step_keep1"][0]),
                                          t(g["step_noise"][0]))
    assert err(mel, g["step_mel"]) < TOL and err(stop, g["step_stop"]) < TOL and err(al, g["step_alignment"]) < TOL
    got = np.stack([st[0][0].numpy(), st[0][1].numpy(), st[1][0].numpy(), st[1][1].numpy()])
    assert err(got, g["step_states"]) < TOL


@pytest.mark.parametrize("name", NAMES)
def test_decoder_loop_matches_reference_sources(name):
    g, cfg, W = load(name)
    out = O.decoder_loop(W, cfg, g["tf_enc"], mels=g["tf_mels"], training=True, keep0=g["tf_keep0"], keep1=g["tf_keep1"],
                         noise=g["tf_noise"])
    assert out["decodings"].shape == g["tf_decodings"].shape
    assert err(out["decodings"], g["tf_decodings"]) < TOL
    assert err(out["stops"], g["tf_stops"]) < TOL
    assert err(out["alignments"], g["tf_alignments"]) < TOL
    fr = O.decoder_loop(W, cfg, g["tf_enc"], training=False, keep0=g["fr_keep0"], keep1=g["fr_keep1"], noise=g["fr_noise"])
    assert fr["decodings"].shape == g["fr_decodings"].shape  # Max_Step // r steps of r frames (Taco2.py:210-214)
    assert err(fr["decodings"], g["fr_decodings"]) < 1e-8
    assert err(fr["stops"], g["fr_stops"]) < 1e-8
    assert err(fr["alignments"], g["fr_alignments"]) < 1e-8


@pytest.mark.parametrize("name", NAMES)
def test_gst_front_end_matches_reference_sources(name):
    g, cfg, W = load(name)
    style, ref, _ = O.style_token_layer(W, cfg, g["gst_mels"], g["gst_lengths"], return_parts=True)
    assert err(ref, g["gst_ref"]) < TOL
    assert err(style, g["gst_style"]) < 1e-7  # Layer_Norm divides by a small std
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    cat = O.gst_concat(t(g["tf_enc"][:, :, cfg.style_size:]), t(g["tf_enc"][:, 0, :cfg.style_size]))
    assert err(cat, g["cat_out"]) == 0.0


@pytest.mark.parametrize("name", NAMES)
def test_multi_head_attention_matches_reference_sources(name):
    g, _, _ = load(name)
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    res, dist = O.multi_head_attention(t(g["mha_Query_kernel"]), t(g["mha_Query_bias"]), t(g["mha_Value_kernel"]),
                                       t(g["mha_Value_bias"]), t(g["mha_gamma"]), t(g["mha_beta"]), 8, t(g["mha_q"]), t(g["mha_v"]))
    assert err(res, g["mha_out"]) < 1e-8 and err(dist, g["mha_dist"]) < TOL


def test_step_form_lsa_matches_sequence_form_reference_layer():
    """The step-form LSA extension, iterated over a query sequence, must reproduce the reference's
    sequence-form LocationSensitiveAttention.call (Layers.py:345-391)."""
    g, cfg, _ = load("sma_r1")
    import copy
    c = copy.deepcopy(cfg)
    c.attention_type, c.lsa_filters, c.lsa_kernel, c.lsa_cumulate, c.lsa_smoothing = "LSA", 4, 5, True, False
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    W = {
        O.DEC + "/Attention/Alignment_Conv/kernel": t(g["lsa_Alignment_Conv_kernel"]),
        O.DEC + "/Attention/Alignment_Conv/bias": t(g["lsa_Alignment_Conv_bias"]),
        O.DEC + "/Attention/Alignment_Dense/kernel": t(g["lsa_Alignment_Dense_kernel"]),
        O.DEC + "/Attention/Alignment_Dense/bias": t(g["lsa_Alignment_Dense_bias"]),
        O.DEC + "/Attention/bias": t(g["lsa_bias"]),
    }
    q = t(g["lsa_q"]) @ t(g["lsa_Query_kernel"]) + t(g["lsa_Query_bias"])
    v = t(g["lsa_v"]) @ t(g["lsa_Value_kernel"]) + t(g["lsa_Value_bias"])
    cum = torch.zeros(q.shape[0], v.shape[1], dtype=torch.float64)
    for s in range(q.shape[1]):
        al = O.lsa_alignment(W, c, q[:, s], v, cum)
        ctx = torch.einsum("bt,bta->ba", al, v)
        cum = cum + al
        assert err(al, g["lsa_alignments"][:, s]) < TOL
        assert err(ctx, g["lsa_contexts"][:, s]) < TOL


@pytest.mark.parametrize("name", NAMES)
def test_postnet_matches_reference_sources(name):
    """Decoder.call's second return value (Taco2.py:230) from the reference's own Postnet Sequential (Taco2.py:130-147)."""
    from gst_tacotron_b200.weights import init_postnet_weights
    g, cfg, _ = load(name)
    WP = init_postnet_weights(cfg, seed=int(g["postnet_seed"]))
    got = O.postnet(WP, cfg, g["fr_decodings"])
    assert got.shape == g["fr_post_decodings"].shape
    assert err(got, g["fr_post_decodings"]) < TOL
    assert err(got, g["fr_decodings"]) > 0.1   # the Postnet term is not negligible next to the tolerance


def test_encoder_oracle_matches_reference_sources():
    """Encoder.call (Taco2.py:47-51) from the reference's own Sequential (Embedding, Conv1D+BN+ReLU x3, Bidirectional LSTM)
    run on the shim (oracle/make_golden.py --encoder) against the CPU oracle."""
    from gst_tacotron_b200.hparams import load_config
    from gst_tacotron_b200.weights import init_encoder_weights
    g = np.load(os.path.join(GOLD, "encoder", "encoder.npz"))
    cfg = load_config()
    WE = init_encoder_weights(cfg, seed=int(g["encoder_seed"]))
    got = O.encoder(WE, cfg, g["tokens"])
    assert got.shape == g["encodings"].shape == (3, 13, 2 * cfg.encoder_rnn_size)
    assert err(got, g["encodings"]) < TOL
    assert float(np.abs(g["encodings"]).max()) > 0.1


@pytest.mark.skipif(not os.path.isdir("/root/reference/Modules"), reason="needs the reference sources (build container only)")
def test_golden_regeneration_is_deterministic(tmp_path):
    """oracle/make_golden.py executes the reference's own sources on the TF shim; with the shim's initialisers seeded, a
    regenerated variant is bit-identical to the committed file (the mha_* / lsa_* kernels used to differ)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = "/root/reference"
    hp = json.load(open(os.path.join(ref, "Hyper_Parameters.json")))
    over = json.load(open(os.path.join(root, "tests", "golden", "sma_r1.hp.json")))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import make_golden as MG
    for k, v in over.items():
        MG._set(hp, k, v)
    json.dump(hp, open(tmp_path / "Hyper_Parameters.json", "w"))
    json.dump(json.load(open(os.path.join(ref, hp["Token_JSON_Path"]))), open(tmp_path / hp["Token_JSON_Path"], "w"))
    out = tmp_path / "sma_r1.npz"
    subprocess.check_call([sys.executable, os.path.join(root, "oracle", "make_golden.py"), "--worker", "sma_r1", str(out)], cwd=tmp_path)
    a, b = np.load(out), np.load(os.path.join(root, "tests", "golden", "sma_r1.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
