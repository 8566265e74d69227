import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.runtime import Engine
from gst_tacotron_b200.weights import init_weights
B, Tv, T = 256, 150, 200
cfg = load_config(precision="bf16")
eng = Engine(cfg, init_weights(cfg, bias_scale=0.05))
rng = np.random.default_rng(0)
text = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.text_dim)).astype(np.float32), device="cuda")
gst = torch.zeros(B, cfg.style_size, device="cuda")
for _ in range(2):
    eng.decode(enc_text=text, gst=gst, steps=T, rng="philox", seed=1, host_outputs=False)
prof = eng.phase_profile().astype(np.float64)[:128] / T
print("us/step", eng.last_kernel_ms() * 1e3 / T)
print("phase C (slot 4):", prof[:, 4].mean())
for i, n in enumerate(["producer wait empty", "MMA wait full", "producer total", "MMA total"]):
    print(n, round(prof[:, 6 + i].mean()), round(prof[:, 6 + i].min()), round(prof[:, 6 + i].max()))
