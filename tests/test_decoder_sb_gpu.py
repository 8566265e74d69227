"""GPU parity of the small-batch latency kernel (csrc/decoder_bf16_sb.cuh: batch <= 16, free running, SMA; mma.sync GEMVs with the
critical LSTM fragments resident in shared memory, one front CTA per utterance, counter hand-overs) against the fp64 CPU oracle
(Modules/Taco2.py:96-120,182-216) and against the batch-256 kernel it replaces for these shapes.  Tolerance 1e-2 (north_star, bf16 mode),
stop-sign rule as in tests/test_decoder_v2_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests.util import BF16_TOL, make_cfg, make_weights, max_abs, oracle_decode, to_np

pytestmark = pytest.mark.gpu
STOP_MARGIN = 1e-2


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
    clear = np.abs(ref["stops"]) > STOP_MARGIN
    assert np.array_equal((to_np(out["stop"]) < 0)[clear], (ref["stops"] < 0)[clear])


@pytest.mark.parametrize("B,Tv,T", [(1, 82, 12), (1, 5, 6), (2, 37, 20), (3, 256, 5), (5, 150, 8), (8, 64, 10), (8, 255, 4), (2, 300, 4),
                                    (9, 40, 10), (12, 150, 6), (16, 82, 12), (16, 256, 4), (17, 30, 4)])
def test_small_batch_kernel_matches_oracle(eng_bf16, B, Tv, T, monkeypatch):
    monkeypatch.delenv("GSTK_DECODER", raising=False)      # default dispatch: batch <= 16, key_time <= 256, free running -> small-batch kernel
                                                           # (one mma n-tile of batch columns up to 8 utterances, two from 9 to 16;
                                                           #  (2, 300, 4) / (17, 30, 4): beyond its limits -> batch-256 kernel)
    cfg, W, eng = eng_bf16
    enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
    ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
    out = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz)
    _check(out, ref)


def test_small_batch_kernel_is_the_one_that_ran_and_agrees_with_the_barrier_kernel(eng_bf16, monkeypatch):
    cfg, W, eng = eng_bf16
    B, Tv, T = 4, 82, 30
    rng = np.random.default_rng(3)
    enc = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
    want = ("mel", "stop", "alignment", "states", "context")
    monkeypatch.delenv("GSTK_DECODER", raising=False)
    a = eng.decode(encodings=enc, steps=T, rng="philox", seed=2, want=want)
    ms_small = eng.last_kernel_ms()
    monkeypatch.setenv("GSTK_DECODER", "barrier")
    b = eng.decode(encodings=enc, steps=T, rng="philox", seed=2, want=want)
    ms_big = eng.last_kernel_ms()
    for k in want:
        assert max_abs(a[k], b[k]) < BF16_TOL, k
    assert not torch.equal(torch.as_tensor(a["mel"]), torch.as_tensor(b["mel"]))   # two different kernels (rounding points differ)
    assert ms_small < ms_big                                                       # ... and the latency kernel is the faster one


def test_small_batch_state_handover_and_long_decode(eng_bf16, monkeypatch):
    """split decode == one decode (states, alignment, last frame handed over); 300 steps stay finite and normalised"""
    monkeypatch.delenv("GSTK_DECODER", raising=False)
    cfg, W, eng = eng_bf16
    B, Tv, T, h = 2, 44, 14, 6
    rng = np.random.default_rng(5)
    enc = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
    full = eng.decode(encodings=enc, steps=T, rng="philox", seed=9)
    a = eng.decode(encodings=enc, steps=h, rng="philox", seed=9, want=("mel", "stop", "alignment", "states"))
    b = eng.decode(encodings=enc, steps=T - h, rng="philox", seed=9, step_offset=h, init_mel=a["mel"][:, -1].contiguous(),
                   init_alignment=a["alignment"][:, -1].contiguous(), init_states=a["states"])
    for k in ("mel", "stop", "alignment"):
        assert max_abs(torch.cat([a[k], b[k]], 1), full[k]) < 5e-3, k
    long = eng.decode(encodings=enc, steps=300, rng="philox", seed=1)
    al = to_np(long["alignment"])
    assert np.isfinite(to_np(long["mel"])).all() and np.all(al >= -1e-6) and np.all(al.sum(-1) < 1 + 1e-3)
    again = eng.decode(encodings=enc, steps=300, rng="philox", seed=1)
    for k in ("mel", "stop", "alignment"):
        assert torch.equal(torch.as_tensor(long[k]), torch.as_tensor(again[k])), k     # bitwise repeatable


def test_small_batch_host_outputs_time_chunks(eng_bf16, monkeypatch):
    """numpy in -> host outputs -> several launches over time with in-place state hand-over (api.cu), same result as one launch"""
    monkeypatch.delenv("GSTK_DECODER", raising=False)
    cfg, W, eng = eng_bf16
    B, Tv, T = 3, 60, 302
    enc = np.random.default_rng(21).uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32)
    a = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
    monkeypatch.setenv("GSTK_NO_TCHUNK", "1")
    b = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
    for k in ("mel", "stop", "alignment"):
        assert max_abs(a[k], b[k]) < 1e-6, k


def test_kernel_selector_of_the_c_abi(eng_bf16, monkeypatch):
    """GstkDecodeArgs::kernel: "batch" at batch 3 = the batch-256 kernel bit for bit (what GSTK_DECODER=barrier selects), rows of a
    small decode are independent of the batch they ran in (what gst_tacotron_b200/shard.py relies on), "small" outside its domain
    is an error, not a silent switch of kernels."""
    cfg, W, eng = eng_bf16
    monkeypatch.delenv("GSTK_DECODER", raising=False)
    B, Tv, T = 3, 40, 9
    enc = torch.as_tensor(np.random.default_rng(5).uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
    pinned = eng.decode(encodings=enc, steps=T, rng="philox", seed=4, kernel="batch")
    small = eng.decode(encodings=enc, steps=T, rng="philox", seed=4, kernel="small")
    auto = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
    monkeypatch.setenv("GSTK_DECODER", "barrier")
    env = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
    monkeypatch.delenv("GSTK_DECODER", raising=False)
    for k in ("mel", "stop", "alignment"):
        assert torch.equal(torch.as_tensor(pinned[k]), torch.as_tensor(env[k])), k
        assert torch.equal(torch.as_tensor(small[k]), torch.as_tensor(auto[k])), k
    one = eng.decode(encodings=enc[1:2], steps=T, rng="philox", seed=4, row_offset=1)      # row 1 alone, same Philox row key
    for k in ("mel", "stop", "alignment"):
        assert torch.equal(torch.as_tensor(one[k])[0], torch.as_tensor(auto[k])[1]), k
    big = torch.zeros(17, Tv, cfg.enc_dim, device="cuda:0")
    with pytest.raises(ValueError, match="GSTK_KERNEL_SMALL"):
        eng.decode(encodings=big, steps=2, kernel="small")
    with pytest.raises(KeyError):
        eng.decode(encodings=enc, steps=2, kernel="fastest")


def test_full_max_step_free_running_against_oracle(eng_bf16, monkeypatch):
    """The reference always runs Max_Step // r = 1000 free-running steps (Taco2.py:210-216).  8 utterances x 150 keys x 1000 steps
    through the batch-256 kernel and the small-batch kernel against the fp64 oracle over the WHOLE trip: the feedback loop does not
    amplify the bf16 rounding differences (running maximum 3e-3, tools/drift_vs_oracle.py), and no decidable stop decision differs."""
    cfg, W, eng = eng_bf16
    monkeypatch.delenv("GSTK_DECODER", raising=False)
    B, Tv, T = 8, 150, 1000
    enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
    ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
    for kernel in ("batch", "small"):
        out = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz, kernel=kernel)
        _check(out, ref)
