"""Early stop: the reference burns Max_Step // r steps and the caller cuts every utterance at
np.argmax(stop < 0) (Model.py:380,413).  gstk_decode(early_stop=1) ends the persistent loop once EVERY utterance has produced a
negative stop logit.  Parity property: everything up to the exit step is bit-identical to the full-length decode, and the
returned stop indices are the first negative entries of the full decode's own stop logits."""
import numpy as np
import pytest
import torch

from tests.util import make_cfg, make_weights, to_np

pytestmark = pytest.mark.gpu
STOP_BIAS = "Decoder/Decoder_Step/Projection/bias"


def _first_negative(stop):
    neg = stop < 0
    return np.where(neg.any(1), neg.argmax(1), stop.shape[1]).astype(np.int32)


def _engine(precision, shift=0.0):
    from gst_tacotron_b200.runtime import Engine
    cfg = make_cfg("SMA", precision=precision)
    W = dict(make_weights(cfg))
    b = np.array(W[STOP_BIAS], np.float32, copy=True)
    b[-1] += shift          # the stop logit is not fed back: shifting its bias moves the stop frames, not the trajectory
    W[STOP_BIAS] = b
    return cfg, W, Engine(cfg, W)


@pytest.mark.parametrize("precision,kernel,B,Tv,T", [("bf16", "barrier", 8, 40, 60), ("bf16", "barrier", 200, 33, 40), ("bf16", "barrier", 300, 24, 24),
                                                   ("bf16", "dataflow", 8, 40, 30), ("fp32", "barrier", 5, 30, 16)])
def test_early_stop_prefix_is_identical_to_full_decode(precision, kernel, B, Tv, T, monkeypatch):
    monkeypatch.setenv("GSTK_DECODER", kernel)
    cfg, W, eng = _engine(precision)
    try:
        rng = np.random.default_rng(17)
        enc = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32), device="cuda:0")
        full = eng.decode(encodings=enc, steps=T, rng="philox", seed=5, want=("mel", "stop", "alignment", "stop_index"))
        stop = to_np(full["stop"])
        # choose a bias shift that makes the LAST utterance stop around 2/3 of the decode (per-row minimum of the logits)
        shift = -float(np.sort(stop[:, : (2 * T) // 3].min(1))[-1]) - 1e-3
        assert np.array_equal(full["stop_index"], _first_negative(stop)) and full["steps_done"] == T
    finally:
        eng.close()
    cfg, W, eng = _engine(precision, shift)
    try:
        full = eng.decode(encodings=enc, steps=T, rng="philox", seed=5, want=("mel", "stop", "alignment", "stop_index"))
        idx = _first_negative(to_np(full["stop"]))
        assert idx.max() < T, "the chosen shift must stop every utterance inside the decode"
        assert np.array_equal(full["stop_index"], idx)
        es = eng.decode(encodings=enc, steps=T, rng="philox", seed=5, early_stop=True)
        assert np.array_equal(es["stop_index"], idx)
        n = es["steps_done"]
        in_kernel = precision == "bf16" and kernel == "barrier"      # the production kernel leaves the loop; the others run all steps
        # per 256-row launch the loop ends right after the step that produced the last negative stop logit
        ends = [(b0, min(b0 + 256, B), (int(idx[b0:b0 + 256].max()) + 1) if in_kernel else T) for b0 in range(0, B, 256)]
        assert n == max(e for _, _, e in ends)
        for k in ("mel", "stop", "alignment"):
            got, ref = torch.as_tensor(es[k]).cpu(), torch.as_tensor(full[k]).cpu()
            assert got.shape[1] == n
            for b0, b1, e in ends:     # every 256-row launch stops on its own; beyond its exit the rows read as zero
                assert torch.equal(got[b0:b1, :e], ref[b0:b1, :e]), k
                # (the exit step itself still runs its attention: one more alignment row is written)
                assert not got[b0:b1, e + (1 if k == "alignment" else 0):].any(), k
    finally:
        eng.close()


def test_early_stop_skips_time_chunks_of_a_host_output_decode(monkeypatch):
    """Host output buffers + >= 256 steps run as several launches over time (api.cu): the launches after the exit are skipped."""
    monkeypatch.setenv("GSTK_DECODER", "barrier")
    cfg, W, eng = _engine("bf16")
    B, Tv, T = 37, 60, 302
    rng = np.random.default_rng(21)
    enc = rng.uniform(-1, 1, (B, Tv, cfg.enc_dim)).astype(np.float32)
    try:
        full = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
        stop = to_np(full["stop"])
        shift = -float(np.sort(stop[:, :100].min(1))[-1]) - 1e-3       # last utterance stops inside the second time chunk at the latest
    finally:
        eng.close()
    cfg, W, eng = _engine("bf16", shift)
    try:
        full = eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
        idx = _first_negative(to_np(full["stop"]))
        k0 = eng.launch_count
        es = eng.decode(encodings=enc, steps=T, rng="philox", seed=4, early_stop=True)
        launched = eng.launch_count - k0
        n = es["steps_done"]
        assert n == int(idx.max()) + 1 and n <= 100 and np.array_equal(es["stop_index"], idx)
        for k in ("mel", "stop", "alignment"):
            assert np.array_equal(to_np(es[k]), to_np(full[k])[:, :n]), k
        k1 = eng.launch_count
        eng.decode(encodings=enc, steps=T, rng="philox", seed=4)
        assert launched < eng.launch_count - k1 + 2, "the early-stopped decode must not launch more decoder kernels than the full one"
    finally:
        eng.close()


def test_early_stop_rejected_for_teacher_forced():
    cfg, W, eng = _engine("bf16")
    try:
        enc = np.zeros((2, 10, cfg.enc_dim), np.float32)
        with pytest.raises(ValueError):
            eng.decode(encodings=enc, teacher_mels=np.zeros((2, 4, cfg.mel_dim), np.float32), early_stop=True)
    finally:
        eng.close()
