"""Kernel time of the bf16 decoder per step for a few shapes, dataflow kernel vs barrier kernel (GSTK_DECODER=dataflow | barrier).
usage: python tools/bench_decoder.py [B,Tv,T ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.runtime import Engine
from gst_tacotron_b200.weights import init_weights

shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(256, 150, 300), (64, 150, 300), (8, 150, 300), (4, 82, 300), (1, 82, 300)]
cfg = load_config(precision="bf16")
eng = Engine(cfg, init_weights(cfg, bias_scale=0.05))
rng = np.random.default_rng(0)
for B, Tv, T in shapes:
    text = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.text_dim)).astype(np.float32), device="cuda")
    gst = torch.zeros(B, cfg.style_size, device="cuda")
    row = []
    for which in ("dataflow", "barrier", None):      # None = default dispatch (small-batch kernel at batch <= 8)
        if which is None:
            os.environ.pop("GSTK_DECODER", None)
        else:
            os.environ["GSTK_DECODER"] = which
        ms = []
        for _ in range(4):
            eng.decode(enc_text=text, gst=gst, steps=T, rng="philox", seed=1, host_outputs=False)
            ms.append(eng.last_kernel_ms())
        row.append(min(ms[1:]) * 1e3 / T)
    print("B={:4d} Tv={:4d} T={:4d}: dataflow {:7.2f} us/step   barrier {:7.2f} us/step   default dispatch {:7.2f} us/step".format(B, Tv, T, *row))
