"""Device time of the wav side at BASELINE's batch shape: Vocoder_Taco1 on [B, T, 80] Postnet output, Griffin-Lim (60 iterations) on
its [B, T, 513] spectrogram.   python tools/bench_vocoder.py [B] [T] [precision]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gst_tacotron_b200.hparams import load_config  # noqa: E402
from gst_tacotron_b200.runtime import Engine  # noqa: E402
from gst_tacotron_b200.weights import init_vocoder_weights, init_weights  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    reps = int(os.environ.get("REPS", "5"))
    cfg = load_config(precision=prec)
    W = dict(init_weights(cfg, seed=1))
    W.update(init_vocoder_weights(cfg))
    eng = Engine(cfg, W)
    g = torch.Generator(device="cuda").manual_seed(0)
    mels = torch.randn(B, T, cfg.mel_dim, device="cuda", generator=g) * 1.5
    ms = []
    for i in range(reps + 2):
        spec = eng.vocoder(mels)
        ms.append(eng.last_kernel_ms())
    voc = float(np.median(ms[2:]))
    # per frame: conv bank 2*8*80*2048, projections 2*3*2048*128 + 2*3*128*128, dense/highway 2*(128*80+80*128+4*128*256),
    # LSTM 2*(128+256)*2048, Dense 2*512*513
    flop = B * T * 2.0 * (8 * 80 * 2048 + 3 * 2048 * 128 + 3 * 128 * 128 + 128 * 80 + 80 * 128 + 4 * 128 * 256 + (128 + 256) * 2048 + 512 * 513)
    print("vocoder  B=%d T=%d %s: %.3f ms  (%.1f M frames/s, %.1f TFLOP/s algorithmic)" % (B, T, prec, voc, B * T / voc / 1e3, flop / voc / 1e9))
    lengths = torch.full((B,), T, dtype=torch.int32, device="cuda")
    for iters in (0, cfg.griffin_lim_iters):
        ms = []
        for i in range(max(2, reps // 2) + 1):
            wav = eng.griffin_lim(spec, lengths=lengths, iters=iters, rng="philox", seed=i, max_abs_value=4.0)
            ms.append(eng.last_kernel_ms())
        gl = float(np.median(ms[1:]))
        print("griffin-lim %2d iterations: %.3f ms  (%.2f ms per iteration, %.1f M samples/s)" % (
            iters, gl, gl / (iters + 1), B * wav.shape[1] / gl / 1e3))
    assert torch.isfinite(wav).all()


if __name__ == "__main__":
    main()
