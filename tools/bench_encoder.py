"""Text Encoder (Taco2.py:12-51) timing at BASELINE configs[2] size: CUDA events around gstk_encoder, device-resident tokens.
usage: python tools/bench_encoder.py [B] [Tv] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gst_tacotron_b200.hparams import load_config  # noqa: E402
from gst_tacotron_b200.runtime import Engine  # noqa: E402
from gst_tacotron_b200.weights import init_encoder_weights, init_weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Tv = int(sys.argv[2]) if len(sys.argv) > 2 else 150
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for precision in ("bf16", "fp32"):
    cfg = load_config()
    cfg.precision = precision
    W = dict(init_weights(cfg))
    W.update(init_encoder_weights(cfg))
    eng = Engine(cfg, W)
    tokens = torch.randint(0, cfg.vocab_size, (B, Tv), device="cuda", dtype=torch.int32)
    ms = []
    for _ in range(reps + 2):
        eng.encoder(tokens)
        ms.append(eng.last_kernel_ms())
    ms = sorted(ms[2:])
    print("encoder {} B={} Tv={}: {:.3f} ms median ({:.3f} min), {:.2f} M tokens/s".format(
        precision, B, Tv, ms[len(ms) // 2], ms[0], B * Tv / ms[len(ms) // 2] * 1e-3))
    eng.close()
