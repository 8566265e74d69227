"""Per-phase time of the small-batch decoder kernel (decoder_bf16_sb.cuh; in-kernel clock64 counters, diagnostics).
usage: python tools/profile_phases_sb.py [B] [Tv] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GSTK_DEBUG"] = str(int(os.environ.get("GSTK_DEBUG", "0")) | 8)
os.environ.pop("GSTK_DECODER", None)

import numpy as np
import torch

from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.runtime import Engine
from gst_tacotron_b200.weights import init_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
Tv = int(sys.argv[2]) if len(sys.argv) > 2 else 82
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
mhz = 1965
print("B={} Tv={} T={}: kernel {:.3f} ms = {:.2f} us/step".format(B, Tv, T, ms, ms * 1e3 / T))
lstm = [(0, "wait h2(t-1) + load"), (1, "U2 / U1 products (fragments from L2)"), (2, "wait [p||ctx](t)"), (3, "load x + W1x GEMV + LSTMCell 0 + publish h1"),
        (4, "wait h1(t)"), (5, "load h1 + W2 GEMV + LSTMCell 1 + publish h2")]
front = [(8, "wait h2(t-1) (+ draws)"), (9, "load + projection + outputs"), (12, "prenet layer 0"), (13, "prenet layer 1"), (10, "query layer"),
         (14, "attention: energies"), (15, "attention: alignment recurrence"), (11, "attention: context + publish x")]
for name, rows, slots in (("LSTM CTAs", list(range(128)), lstm), ("front CTAs", list(range(128, 128 + B)), front)):
    print(" ", name)
    for i, n in slots:
        col = prof[rows, i] / T
        print("    {:<52s} mean {:7.0f} ticks  min {:7.0f}  max {:7.0f}  (~{:5.2f} us)".format(n, col.mean(), col.min(), col.max(), col.mean() / mhz))
