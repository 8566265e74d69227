"""BASELINE.json configurations at their full sizes, checked through size-independent properties (the CPU oracle cannot
run 64 x 800 or 256 x 1000 steps in test time):
  configs[1]  teacher-forced batch=64, 150 tokens x 800 mel frames          (bf16 + fp32 engines)
  configs[2]  free-running batch=256, stepwise-monotonic vs location-sensitive attention, Max_Step cap
Properties: one launch == two launches with state hand-over (the recurrence is a pure function of the carried state),
the first steps of the long decode == the oracle on those steps (prefix property of a causal loop), alignment rows are
sub-stochastic and monotone in support (SMA), everything finite, bf16 and fp32 engines agree on clear stop decisions."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests.util import BF16_TOL, FP32_TOL, make_cfg, make_weights, max_abs, oracle_decode, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engs():
    from gst_tacotron_b200.runtime import Engine
    out = {}
    for prec in ("bf16", "fp32"):
        cfg = make_cfg("SMA", precision=prec)
        W = make_weights(cfg)
        out[prec] = (cfg, W, Engine(cfg, W))
    yield out
    for _, _, e in out.values():
        e.close()


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_config1_teacher_forced_64x150x800(engs, prec):
    cfg, W, eng = engs[prec]
    B, Tv, T = 64, 150, 800
    rng = np.random.default_rng(2024)
    enc = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
    mels = rng.uniform(-4, 4, (B, T + 1, cfg.mel_dim)).astype(np.float32)
    mels[:, 0] = 0
    teach = torch.as_tensor(mels[:, :-1], device="cuda:0")
    full = eng.decode(encodings=enc, teacher_mels=teach, rng="philox", seed=7, want=("mel", "stop", "alignment", "states"))
    al = to_np(full["alignment"])
    assert al.shape == (B, T, Tv) and np.isfinite(al).all() and np.isfinite(to_np(full["mel"])).all()
    assert (al >= 0).all() and (al.sum(-1) <= 1 + 1e-3).all()          # SMA rows are sub-stochastic (mass leaks at the end)
    reach = np.broadcast_to(np.arange(Tv)[None, None, :] > (np.arange(T)[None, :, None] + 1), al.shape)
    assert not al[reach].any()                                           # alignment can move one key per step: support <= t + 1
    # causal prefix == oracle on the first steps
    h = 6
    k0, k1, nz = O.philox_randomness(cfg, 7, h, B, Tv)
    ref = oracle_decode(cfg, W, to_np(enc), mels=mels[:, : h + 1], keep0=k0, keep1=k1, noise=nz, dtype=torch.float32)
    tol = BF16_TOL if prec == "bf16" else FP32_TOL
    assert max_abs(to_np(full["mel"])[:, :h], ref["decodings"]) < tol
    assert max_abs(al[:, :h], ref["alignments"]) < tol
    # one launch == two launches with state hand-over
    s = 333
    a = eng.decode(encodings=enc, teacher_mels=teach[:, :s].contiguous(), rng="philox", seed=7, want=("mel", "stop", "alignment", "states"))
    b = eng.decode(encodings=enc, teacher_mels=teach[:, s:].contiguous(), rng="philox", seed=7, step_offset=s,
                   init_alignment=a["alignment"][:, -1].contiguous(), init_states=a["states"])
    assert max_abs(torch.cat([torch.as_tensor(a["mel"]), torch.as_tensor(b["mel"])], 1), full["mel"]) < 1e-5
    assert max_abs(torch.cat([torch.as_tensor(a["stop"]), torch.as_tensor(b["stop"])], 1), full["stop"]) < 1e-5


def test_config2_free_running_256_sma_vs_lsa_max_step(engs):
    from gst_tacotron_b200.runtime import Engine
    B, Tv = 256, 150
    rng = np.random.default_rng(5)
    outs = {}
    for att in ("SMA", "LSA"):
        cfg = make_cfg(att, precision="bf16")
        T = cfg.max_step // cfg.step_reduction   # the reference always runs Max_Step // r steps (Taco2.py:210-214)
        W = make_weights(cfg)
        eng = engs["bf16"][2] if att == "SMA" else Engine(cfg, W)
        enc = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
        out = eng.decode(encodings=enc, steps=T, rng="philox", seed=1, host_outputs=False)
        mel, stop, al = (torch.as_tensor(out[k]) for k in ("mel", "stop", "alignment"))
        assert mel.shape == (B, T * cfg.step_reduction, cfg.mel_dim) and stop.shape == (B, T) and al.shape == (B, T, Tv)
        assert bool(torch.isfinite(mel).all()) and bool(torch.isfinite(stop).all()) and bool(torch.isfinite(al).all())
        if att == "LSA":
            assert float((al.sum(-1) - 1).abs().max()) < 1e-3            # softmax rows (Layers.py:419)
            eng.close()
        else:
            assert float(al.sum(-1).max()) <= 1 + 1e-3
        outs[att] = out


@pytest.mark.parametrize("n_tokens", [16, 10])
def test_config3_gst_512x1000(n_tokens):
    """configs[3]: reference encoder (6 conv layers + GRU) + 4-head style attention over batch=512 x 1000-frame mels.
    A handful of rows is checked against the oracle, the rest through batch independence (each row only depends on its own
    mel and length) and the Layer_Norm invariant."""
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA", n_tokens=n_tokens)
    W = make_weights(cfg)
    eng = Engine(cfg, W)
    B, T = 512, 1000
    rng = np.random.default_rng(4)
    mels = rng.uniform(-4, 4, (B, T + 1, cfg.mel_dim)).astype(np.float32)
    mels[:, 0] = 0
    lens = rng.integers(640, T + 1, B).astype(np.int32)
    md = torch.as_tensor(mels, device="cuda:0")
    out = eng.gst(md, lens, drop_first=True, want=("gst",))
    g = to_np(out["gst"])
    assert g.shape == (B, cfg.style_size) and np.isfinite(g).all()
    assert np.allclose(g.mean(-1), 0.0, atol=1e-4)
    pick = np.array([0, 17, 255, 511])
    ref = O.style_token_layer(W, cfg, mels[pick], lens[pick])
    assert max_abs(g[pick], ref) < FP32_TOL
    sub = eng.gst(md[100:164].contiguous(), lens[100:164], drop_first=True, want=("gst",))
    assert max_abs(sub["gst"], g[100:164]) < 1e-6
    eng.close()
