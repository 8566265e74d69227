"""GPU parity of BOTH bf16 free-running decoder kernels against the fp64 CPU oracle (Modules/Taco2.py:96-120,182-216):
the production kernel (csrc/decoder_bf16.cuh, grid-barrier phases) and the opt-in dataflow variant (csrc/decoder_bf16_v2.cuh,
GSTK_DECODER=dataflow: folded projection -> prenet-0 matrix on the tensor cores, arrival counters instead of grid barriers).
Includes the exact configuration bench.py times (batch 256, 150 keys, free running, Philox randomness).

Tolerance: 1e-2 absolute on mel / stop / alignment (north_star, bf16 mode).  Free-running stop-frame rule:
on the compared prefix the sign of the stop logit must be identical wherever |oracle stop| > STOP_MARGIN."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests.util import BF16_TOL, make_cfg, make_weights, max_abs, oracle_decode, to_np

pytestmark = pytest.mark.gpu
STOP_MARGIN = 1e-2


@pytest.fixture(scope="module")
def _engine():
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA", precision="bf16")
    W = make_weights(cfg)
    e = Engine(cfg, W)
    yield cfg, W, e
    e.close()


@pytest.fixture(params=["barrier", "dataflow"])
def eng_bf16(request, _engine, monkeypatch):
    """the same engine, with the decoder kernel selected per test (api.cu reads GSTK_DECODER at every call)"""
    monkeypatch.setenv("GSTK_DECODER", request.param)
    return _engine


def _check(out, ref, tol=BF16_TOL):
    assert np.isfinite(to_np(out["mel"])).all()
    assert max_abs(out["mel"], ref["decodings"]) < tol
    assert max_abs(out["stop"], ref["stops"]) < tol
    assert max_abs(out["alignment"], ref["alignments"]) < tol
    clear = np.abs(ref["stops"]) > STOP_MARGIN
    assert np.array_equal((to_np(out["stop"]) < 0)[clear], (ref["stops"] < 0)[clear])


# ragged batch sizes (row halves / m-tiles partly filled, one dense CTA per few utterances), key_time beyond one attention sweep
@pytest.mark.parametrize("B,Tv,T", [(1, 82, 12), (3, 37, 20), (64, 50, 8), (65, 33, 6), (130, 40, 6), (200, 170, 5), (256, 30, 5), (2, 330, 4)])
def test_v2_free_running_external_randomness_matches_oracle(eng_bf16, B, Tv, T):
    cfg, W, eng = eng_bf16
    enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
    ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
    _check(out, ref)


def test_v2_bench_configuration_prefix_matches_oracle(eng_bf16):
    """The configuration bench.py times (BASELINE configs[2]): bf16, batch 256, 150 keys, free running, Philox randomness.
    First 16 steps against the fp64 oracle fed the same Philox streams."""
    cfg, W, eng = eng_bf16
    B, Tv, T = 256, 150, 16
    rng = np.random.default_rng(7)
    enc = rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32)
    k0, k1, nz = O.philox_randomness(cfg, 13, T, B, Tv)
    ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz, dtype=torch.float32)
    out = eng.decode(encodings=enc, steps=T, rng="philox", seed=13)
    _check(out, ref)
    # the same prefix out of the full-length decode of the benchmark (1000 steps, device tensors, chunk-free)
    enc_d = torch.as_tensor(enc, device="cuda:0")
    full = eng.decode(encodings=enc_d, steps=1000, rng="philox", seed=13, host_outputs=False)
    for k in ("mel", "stop", "alignment"):
        assert torch.equal(torch.as_tensor(full[k])[:, :T].cpu(), torch.as_tensor(out[k]).cpu()), k
    al = to_np(full["alignment"])
    assert np.isfinite(to_np(full["mel"])).all() and np.all(al >= -1e-6) and np.all(al.sum(-1) < 1 + 1e-3)


def test_v2_agrees_with_barrier_kernel(_engine, monkeypatch):
    """Same decode through both kernels: they differ only in bf16 rounding points (the folded matrix skips the bf16 rounding of
    the fed-back frame), far inside the 1e-2 budget over a short horizon."""
    cfg, W, eng = _engine
    monkeypatch.setenv("GSTK_DECODER", "dataflow")
    B, Tv, T = 140, 61, 12
    rng = np.random.default_rng(3)
    enc = rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32)
    a = eng.decode(encodings=enc, steps=T, rng="philox", seed=2, want=("mel", "stop", "alignment", "states", "context"))
    monkeypatch.setenv("GSTK_DECODER", "barrier")
    b = eng.decode(encodings=enc, steps=T, rng="philox", seed=2, want=("mel", "stop", "alignment", "states", "context"))
    for k in ("mel", "stop", "alignment", "states", "context"):
        assert max_abs(a[k], b[k]) < BF16_TOL, k


def test_v2_state_handover_continues_the_decode(eng_bf16):
    """Free-running decode split in two calls (states, alignment and last frame handed over) == one call."""
    cfg, W, eng = eng_bf16
    B, Tv, T, h = 9, 44, 14, 6
    rng = np.random.default_rng(5)
    enc = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
    full = eng.decode(encodings=enc, steps=T, rng="philox", seed=9)
    a = eng.decode(encodings=enc, steps=h, rng="philox", seed=9, want=("mel", "stop", "alignment", "states"))
    b = eng.decode(encodings=enc, steps=T - h, rng="philox", seed=9, step_offset=h, init_mel=a["mel"][:, -1].contiguous(),
                   init_alignment=a["alignment"][:, -1].contiguous(), init_states=a["states"])
    # the hand-over re-enters through the fp32 prenet-0 kernel on the last frame instead of the folded bf16 matrix: equal up to
    # that one rounding difference
    for k in ("mel", "stop", "alignment"):
        assert max_abs(torch.cat([a[k], b[k]], 1), full[k]) < 5e-3, k


def test_v2_lsa_and_teacher_forced_stay_on_the_barrier_kernel(eng_bf16):
    """Dispatch: teacher-forced decodes do not take the folded path (the fed frame is not the projection); result still within
    tolerance of the oracle."""
    cfg, W, eng = eng_bf16
    B, Tv, T = 5, 29, 7
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz)
    assert max_abs(out["mel"], ref["decodings"]) < BF16_TOL and max_abs(out["alignment"], ref["alignments"]) < BF16_TOL


@pytest.mark.parametrize("kernel", ["barrier", "dataflow"])
def test_reload_weights_rebuilds_bf16_images(kernel, monkeypatch):
    """a second load_weights() on a bf16 handle that has already decoded must rebuild the packed LSTM / dense /
    folded images - decode with the NEW weights matches the oracle on the new weights."""
    from gst_tacotron_b200.runtime import Engine
    monkeypatch.setenv("GSTK_DECODER", kernel)
    cfg = make_cfg("SMA", precision="bf16")
    W1, W2 = make_weights(cfg, seed=1), make_weights(cfg, seed=2)
    eng = Engine(cfg, W1)
    try:
        B, Tv, T = 4, 30, 6
        enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
        eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
        enc2, mels2, _, _, _ = O.synth_decoder_inputs(cfg, B, Tv, T)
        eng.decode(encodings=enc2, teacher_mels=mels2[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz)
        eng.load_weights(W2)
        ref = oracle_decode(cfg, W2, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
        out = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
        _check(out, ref)
        ref_t = oracle_decode(cfg, W2, enc2, mels=mels2, keep0=k0, keep1=k1, noise=nz)
        out_t = eng.decode(encodings=enc2, teacher_mels=mels2[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz)
        assert max_abs(out_t["mel"], ref_t["decodings"]) < BF16_TOL
    finally:
        eng.close()


@pytest.mark.parametrize("att", ["LSA", "BMA"])
def test_bench_configuration_prefix_other_attention_types(att):
    """The benchmark shape (batch 256, 150 keys, free running, Philox) with the location-sensitive and the Bahdanau-monotonic
    attention of Hyper_Parameters.json's other `Attention.Type` values: first 16 steps against the oracle (generic phase A of the
    bf16 kernel + the tcgen05 LSTM phases)."""
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg(att, precision="bf16")
    W = make_weights(cfg)
    eng = Engine(cfg, W)
    try:
        B, Tv, T = 256, 150, 16
        enc = np.random.default_rng(8).uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32)
        k0, k1, nz = O.philox_randomness(cfg, 21, T, B, Tv)
        ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz, dtype=torch.float32)
        out = eng.decode(encodings=enc, steps=T, rng="philox", seed=21)
        _check(out, ref)
    finally:
        eng.close()
