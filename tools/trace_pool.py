"""Host- and device-side timeline of two pooled engines on the bench workload (GSTK_TRACE=1 prints the library's trace points on stderr).
   GSTK_TRACE=1 python tools/trace_pool.py 2> trace.txt"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gst_tacotron_b200.hparams import load_config
from gst_tacotron_b200.runtime import EnginePool
from gst_tacotron_b200.weights import init_weights
cfg = load_config(precision="bf16"); W = init_weights(cfg, seed=1)
B, Tv, T = 256, 150, 1000
rng = np.random.default_rng(0)
text_h = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.text_dim)).astype(np.float32)).pin_memory()
mels_h = torch.as_tensor(rng.uniform(-4, 4, (B, 189, cfg.mel_dim)).astype(np.float32)).pin_memory()
lens_h = torch.full((B,), 188, dtype=torch.int32).pin_memory()
pool = EnginePool(cfg, W, depth=2)
outs = [{"mel": torch.empty(B, T, cfg.mel_dim).pin_memory(), "stop": torch.empty(B, T).pin_memory(), "alignment": torch.empty(B, T, Tv).pin_memory()} for _ in range(2)]
def call(e, i, k):
    g = e.gst(mels_h, lens_h, want=("gst",), host_outputs=False)["gst"]
    e.decode(enc_text=text_h, gst=g, steps=T, rng="philox", seed=i, host_outputs=True, want=("mel", "stop", "alignment"), out_buffers=outs[k])
pend = []
def run(n):
    for i in range(n):
        k = i % 2
        if len(pend) >= 2: pend.pop(0).result()
        pend.append(pool.submit(lambda e, i=i, k=k: call(e, i, k), engine=k))
    while pend: pend.pop(0).result()
run(4)
print("=== traced ===", file=sys.stderr, flush=True)
run(6)
pool.close()
