"""GPU parity of the bf16 tensor-core (tcgen05) decoder against the fp64 CPU oracle.
Tolerance: 1e-2 absolute (north_star, bf16 mode)."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests.util import BF16_TOL, make_cfg, make_weights, max_abs, oracle_decode, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_bf16():
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA", precision="bf16")
    W = make_weights(cfg)
    e = Engine(cfg, W)
    yield cfg, W, e
    e.close()


def _check(out, ref, tol=BF16_TOL):
    assert np.isfinite(to_np(out["mel"])).all()
    assert max_abs(out["mel"], ref["decodings"]) < tol
    assert max_abs(out["stop"], ref["stops"]) < tol
    assert max_abs(out["alignment"], ref["alignments"]) < tol


# (150, 170, 4) and (2, 330, 3): key_time beyond one 32-row sweep per warp (several sweeps in the attention passes)
@pytest.mark.parametrize("B,Tv,T", [(1, 82, 10), (3, 37, 20), (64, 50, 8), (130, 40, 6), (256, 30, 5), (150, 170, 4), (2, 330, 3)])
def test_bf16_teacher_forced_matches_oracle(eng_bf16, B, Tv, T):
    cfg, W, eng = eng_bf16
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz)
    _check(out, ref)


def test_bf16_batch_chunking_over_256(eng_bf16):
    cfg, W, eng = eng_bf16
    B, Tv, T = 300, 24, 4
    enc, mels, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="external", keep0=k0, keep1=k1, noise=nz,
                     want=("mel", "stop", "alignment", "states"))
    _check(out, ref)
    (h1, c1), (h2, c2) = ref["states"]
    st = to_np(out["states"])
    # h is bounded by 1; the cell states are open sums (|c| reaches 3-4 with +-4 teacher frames): 1e-2 relative to that scale
    assert max_abs(st[0], h1.numpy()) < BF16_TOL and max_abs(st[1], c1.numpy()) < 2 * BF16_TOL
    assert max_abs(st[2], h2.numpy()) < BF16_TOL and max_abs(st[3], c2.numpy()) < 2 * BF16_TOL


def test_bf16_long_teacher_forced_stays_within_tolerance(eng_bf16):
    """200 recurrent steps: bf16 operand rounding must not drift past the 1e-2 budget."""
    cfg, W, eng = eng_bf16
    B, Tv, T = 8, 150, 200
    enc, mels, _, _, _ = O.synth_decoder_inputs(cfg, B, Tv, T)
    k0, k1, nz = O.philox_randomness(cfg, 5, T, B, Tv)
    ref = oracle_decode(cfg, W, enc, mels=mels, keep0=k0, keep1=k1, noise=nz, dtype=torch.float32)
    out = eng.decode(encodings=enc, teacher_mels=mels[:, :-1], rng="philox", seed=5)
    _check(out, ref)


def test_bf16_free_running_stop_frames(eng_bf16):
    cfg, W, eng = eng_bf16
    B, Tv, T = 4, 40, 30
    enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
    ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
    _check(out, ref)
    # stop-frame identity wherever the logit is not within rounding distance of zero
    clear = np.abs(ref["stops"]) > BF16_TOL
    assert np.array_equal((to_np(out["stop"]) < 0)[clear], (ref["stops"] < 0)[clear])


def test_bf16_state_roundtrip(eng_bf16):
    cfg, W, eng = eng_bf16
    B, Tv, T = 5, 31, 8
    enc, mels, _, _, _ = O.synth_decoder_inputs(cfg, B, Tv, T)
    dev = "cuda:0"
    enc_d = torch.as_tensor(enc, device=dev)
    teach = torch.as_tensor(mels[:, :-1], device=dev)
    full = eng.decode(encodings=enc_d, teacher_mels=teach, rng="philox", seed=9)
    h = 5
    a = eng.decode(encodings=enc_d, teacher_mels=teach[:, :h].contiguous(), rng="philox", seed=9,
                   want=("mel", "stop", "alignment", "states"))
    b = eng.decode(encodings=enc_d, teacher_mels=teach[:, h:].contiguous(), rng="philox", seed=9, step_offset=h,
                   init_alignment=a["alignment"][:, -1].contiguous(), init_states=a["states"])
    # h is re-quantised to bf16 at the split, exactly as inside one launch
    assert max_abs(torch.cat([a["mel"], b["mel"]], 1), full["mel"]) < 1e-5


def test_bf16_repeatable(eng_bf16):
    """Two free-running decodes of the benchmark shape with the same seed must agree bit for bit: every cross-CTA hand-over
    (operand images, queries, barrier generations) is exercised 300 times per run, so a stale read shows up as a mismatch."""
    cfg, W, eng = eng_bf16
    B, Tv, T = 256, 150, 300
    rng = np.random.default_rng(11)
    enc = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
    outs = [eng.decode(encodings=enc, steps=T, rng="philox", seed=3, host_outputs=False) for _ in range(3)]
    for o in outs[1:]:
        for k in ("mel", "stop", "alignment"):
            assert torch.equal(torch.as_tensor(o[k]), torch.as_tensor(outs[0][k])), k
    assert np.isfinite(to_np(outs[0]["mel"])).all()


def test_bf16_time_chunked_host_outputs_match_single_launch(eng_bf16, monkeypatch):
    """With host output buffers a long decode runs as several launches with in-place state hand-over and overlapped D2H copies
    (api.cu); the result must be the same as the single launch."""
    cfg, W, eng = eng_bf16
    B, Tv, T = 37, 60, 302
    rng = np.random.default_rng(21)
    enc = rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32)
    a = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)                      # numpy in -> host outputs -> chunked
    monkeypatch.setenv("GSTK_NO_TCHUNK", "1")
    b = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
    monkeypatch.delenv("GSTK_NO_TCHUNK")
    for k in ("mel", "stop", "alignment"):
        assert to_np(a[k]).shape == to_np(b[k]).shape
        assert max_abs(a[k], b[k]) < 1e-6, k
