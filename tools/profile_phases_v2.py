"""Per-phase time of the dataflow bf16 decoder (decoder_bf16_v2.cuh; in-kernel clock64 counters, diagnostics).
usage: python tools/profile_phases_v2.py [B] [Tv] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GSTK_DEBUG"] = str(int(os.environ.get("GSTK_DEBUG", "0")) | 8)   # bit 3: in-kernel per-phase timers on
os.environ["GSTK_DECODER"] = "dataflow"

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
mhz = 1965
print("B={} Tv={} T={}: kernel {:.3f} ms = {:.2f} us/step".format(B, Tv, T, ms, ms * 1e3 / T))
lstm = [(0, "C-epilogue tail + fold epilogue (until the z0 wait)"), (1, "wait for z0 (fold CTAs of both m-tiles)"), (5, "z0 rows + prenet-1 + query (mma.sync, weights from L2)"),
        (2, "attention"), (3, "wait d1 + LSTMCell-0 epilogue + publish"), (4, "noise/keep draw + wait d2 + LSTMCell-1 epilogue + publish")]
seg = [(15, "MMA warp: fold stream (9 units)"), (6, "MMA warp: h2.U2 (16)"), (7, "MMA warp: h1.U1 (16)"), (12, "MMA warp: p.W1x (4)"), (13, "MMA warp: ctx.W1x (2)"),
       (14, "MMA warp: h1.W2 (16)"), (8, "act-copy warp, fold: waiting for a free stage"), (9, "act-copy warp, fold: polling h2 counters"),
       (10, "act-copy warp, h1.W2: waiting for a free stage"), (11, "act-copy warp, h1.W2: polling h1 counters")]
nl = min(128, prof.shape[0])
act = [c for c in range(nl) if (c & 1) < (B + 127) // 128]
for name, rows in (("fold CTAs (ug < 22)", [c for c in act if (c >> 1) < 22]), ("other LSTM CTAs", [c for c in act if (c >> 1) >= 22])):
    print(" ", name)
    for i, n in lstm + seg:
        col = prof[rows, i] / T
        print("    {:<62s} mean {:7.0f} ticks  min {:7.0f}  max {:7.0f}  (~{:5.2f} us)".format(n, col.mean(), col.min(), col.max(), col.mean() / mhz))
    tot = prof[rows][:, [i for i, _ in lstm]].sum(1).mean() / T
    print("    sum of the phase slots: {:.0f} ticks = {:.2f} us".format(tot, tot / mhz))
