"""GST front end (Reference_Encoder conv stack + GRU + style-token attention) on BASELINE configs[3]: batch 512 x 1000-frame mels,
16 and 10 tokens, tensor-core mode vs fp32 FFMA kernels.  usage: python tools/bench_gst.py [B] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.runtime import Engine
from gst_tacotron_b200.weights import init_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
GST_MFLOP = 201.4   # per 1000-frame mel
for prec, tokens, tc in (("bf16", 16, "1"), ("bf16", 10, "1"), ("bf16", 16, "0"), ("fp32", 16, "1")):
    os.environ["GSTK_GST_TC"] = tc
    cfg = load_config(precision=prec, n_tokens=tokens)
    eng = Engine(cfg, init_weights(cfg, bias_scale=0.05))
    rng = np.random.default_rng(0)
    mels = torch.as_tensor(rng.uniform(-4, 4, (B, T, cfg.mel_dim)).astype(np.float32), device="cuda")
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    for _ in range(3):
        eng.gst(mels, lens)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        eng.gst(mels, lens)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = B * GST_MFLOP * (T / 1000.0) * 1e6 / (ms * 1e-3) / 1e12
    print("precision={} tokens={} conv_tc={}: {} x {} frames {:.3f} ms = {:.1f} TFLOP/s algorithmic".format(prec, tokens, tc, B, T, ms, tf))
    eng.close()
