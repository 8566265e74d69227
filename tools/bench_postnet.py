"""Postnet (Taco2.py:130-149,:230) timing at BASELINE configs[2] size: CUDA events around gstk_postnet, device-resident
input.  usage: python tools/bench_postnet.py [B] [T] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gst_tacotron_b200.hparams import load_config  # noqa: E402
from gst_tacotron_b200.runtime import Engine  # noqa: E402
from gst_tacotron_b200.weights import init_postnet_weights, init_weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for precision in ("bf16", "fp32"):
    cfg = load_config()
    cfg.precision = precision
    W = dict(init_weights(cfg))
    W.update(init_postnet_weights(cfg))
    eng = Engine(cfg, W)
    dec = torch.rand(B, T, cfg.mel_dim, device="cuda") * 8 - 4
    cin = [cfg.mel_dim] + [l[0] for l in cfg.postnet_layers[:-1]]
    flops = 2.0 * B * T * sum(k * ci * co for (co, k, _s, _t), ci in zip(cfg.postnet_layers, cin))
    ms = []
    for _ in range(reps + 2):
        eng.postnet(dec)
        ms.append(eng.last_kernel_ms())
    ms = sorted(ms[2:])
    med = ms[len(ms) // 2]
    print("postnet {} B={} T={}: {:.3f} ms median ({:.3f} min), {:.1f} TFLOP/s algorithmic, {:.2f} M frames/s".format(
        precision, B, T, med, ms[0], flops / med * 1e-9, B * T / med * 1e-3))
    eng.close()
