"""Per-phase time of the persistent bf16 decoder (in-kernel clock64 counters, diagnostics).
usage: python tools/profile_phases.py [B] [Tv] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
names = ["phase A", "barrier 1", "phase B (LSTM1)", "barrier 2", "phase C (LSTM2)", "barrier 3"]
print("B={} Tv={} T={}: kernel {:.3f} ms = {:.2f} us/step".format(B, Tv, T, ms, ms * 1e3 / T))
tot = prof[:, :6].sum(1).mean()
for i, n in enumerate(names):
    col = prof[:, i] / T
    print("  {:<18s} mean {:8.0f} ticks  min {:8.0f}  max {:8.0f}   (~{:5.2f} us mean at {} MHz)".format(
        n, col.mean(), col.min(), col.max(), col.mean() / mhz, mhz))
print("  sum of phases per step: {:.0f} ticks".format(tot / T))
lstm = prof[:128, :6] / T
rest = prof[128:, :6] / T
sub = ["A: projection(t-1)+out pass", "A: input staging+sync", "A: prenet0 mma+sync", "A: prenet0 act pass+sync", "A: prenet1 mma+sync",
       "A: prenet1 act pass+sync", "A: query mma+sync", "A: query pass+sync", "A: attention (all)", "A: (of which) waiting for weight stages"]
for i, n in enumerate(sub):
    col = prof[:, 6 + i] / T
    print("  {:<30s} mean {:8.0f} ticks  min {:8.0f}  max {:8.0f}".format(n, col.mean(), col.min(), col.max()))
print("  LSTM CTAs  (0-127) mean per phase:", np.round(lstm.mean(0)).astype(int))
print("  other CTAs (128+)  mean per phase:", np.round(rest.mean(0)).astype(int))
