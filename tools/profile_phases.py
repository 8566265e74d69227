"""Per-phase time of the persistent bf16 decoder (in-kernel clock64 counters, diagnostics).
usage: python tools/profile_phases.py [B] [Tv] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GSTK_DEBUG"] = str(int(os.environ.get("GSTK_DEBUG", "0")) | 8)   # bit 3: in-kernel per-phase timers on

import numpy as np
import torch

from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.runtime import Engine
from gst_tacotron_b200.weights import init_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Tv = int(sys.argv[2]) if len(sys.argv) > 2 else 150
T = int(sys.argv[3]) if len(sys.argv) > 3 else 200
cfg = load_config(precision="bf16")
eng = Engine(cfg, init_weights(cfg, bias_scale=0.05))
rng = np.random.default_rng(0)
text = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.text_dim)).astype(np.float32), device="cuda")
gst = torch.zeros(B, cfg.style_size, device="cuda")
for _ in range(2):
    eng.decode(enc_text=text, gst=gst, steps=T, rng="philox", seed=1, host_outputs=False)
ms = eng.last_kernel_ms()
prof = eng.phase_profile().astype(np.float64)
mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
names = ["A1 dense layers (rest)", "barrier 0", "A2 attention", "barrier 1", "phase B (LSTM1)", "hand-over B->C (flags)", "phase C (LSTM2)", "barrier 3"]
NP = len(names)
print("B={} Tv={} T={}: kernel {:.3f} ms = {:.2f} us/step".format(B, Tv, T, ms, ms * 1e3 / T))
for i, n in enumerate(names):
    col = prof[:, i] / T
    print("  {:<24s} mean {:8.0f} ticks  min {:8.0f}  max {:8.0f}   (~{:5.2f} us mean at {} MHz)".format(
        n, col.mean(), col.min(), col.max(), col.mean() / mhz, mhz))
print("  sum of all slots per step (mean over CTAs): {:.0f} ticks".format(prof.sum(1).mean() / T))
sub = ["A1: [h2|ctx] rows load+sync", "A1: projection mma+sync", "A1: out pass+input staging", "A1: prenet0 mma+epilogue+sync", "A1: (warp 0) waiting for weight stages",
       "A1: prenet1 mma+epilogue+sync", "(unused)", "A1: query mma+epilogue"]
dense = prof[128:]
for i, n in enumerate(sub):
    col = dense[:, 8 + i] / T
    print("  {:<30s} (dense CTAs) mean {:8.0f} ticks  min {:8.0f}  max {:8.0f}".format(n, col.mean(), col.min(), col.max()))
print("  LSTM CTAs  (0-127) mean per slot:", np.round(prof[:128, :NP].mean(0) / T).astype(int))
print("  dense CTAs (128+)  mean per slot:", np.round(dense[:, :NP].mean(0) / T).astype(int))
