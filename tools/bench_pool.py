"""Where the end-to-end step goes: the bench workload (256 utterances x 150 keys x 1000 steps + GST) through Engine / EnginePool with
host or device buffers on either side.   python tools/bench_pool.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gst_tacotron_b200.hparams import load_config  # noqa: E402
from gst_tacotron_b200.runtime import EnginePool  # noqa: E402
from gst_tacotron_b200.weights import init_weights  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    cfg = load_config(precision="bf16")
    W = init_weights(cfg, seed=1)
    B, Tv, T = 256, 150, 1000
    rng = np.random.default_rng(0)
    text_h = torch.as_tensor(rng.uniform(-1, 1, (B, Tv, cfg.text_dim)).astype(np.float32)).pin_memory()
    gst_h = torch.as_tensor(rng.uniform(-1, 1, (B, cfg.style_size)).astype(np.float32)).pin_memory()
    text_d, gst_d = text_h.cuda(), gst_h.cuda()
    mels_h = torch.as_tensor(rng.uniform(-4, 4, (B, 189, cfg.mel_dim)).astype(np.float32)).pin_memory()
    lens_h = torch.full((B,), 188, dtype=torch.int32).pin_memory()
    mels_d, lens_d = mels_h.cuda(), lens_h.cuda()
    for depth in [int(d) for d in os.environ.get("DEPTHS", "1,2").split(",")]:
        pool = EnginePool(cfg, W, depth=depth)
        outs = [{"mel": torch.empty(B, T, cfg.mel_dim).pin_memory(), "stop": torch.empty(B, T).pin_memory(),
                 "alignment": torch.empty(B, T, Tv).pin_memory()} for _ in range(depth)]
        for name, host_in, want in (("host in, host out (mel+stop+alignment)", True, ("mel", "stop", "alignment")),
                                    ("host in, host out (mel+stop)", True, ("mel", "stop")),
                                    ("device in, host out (mel+stop+alignment)", False, ("mel", "stop", "alignment")),
                                    ("host in, device out", True, None),
                                    ("device in, device out", False, None),
                                    ("GST(host mels) + host in, host out (all)", "gst", ("mel", "stop", "alignment")),
                                    ("GST(device mels) + host in, host out (all)", "gstd", ("mel", "stop", "alignment")),
                                    ("GST(device mels) + device in, device out", "gstdd", None)):
            def call(e, i, k):
                g = gst_h if host_in is True else gst_d
                if host_in == "gst":
                    g = e.gst(mels_h, lens_h, want=("gst",), host_outputs=False)["gst"]
                elif host_in in ("gstd", "gstdd"):
                    g = e.gst(mels_d, lens_d, want=("gst",), host_outputs=False)["gst"]
                kw = dict(enc_text=text_d if host_in in (False, "gstdd") else text_h, gst=g, steps=T, rng="philox", seed=i)
                if want is None:
                    e.decode(host_outputs=False, want=("mel", "stop", "alignment"), **kw)
                    e.synchronize()
                else:
                    e.decode(host_outputs=True, want=want, out_buffers={w: outs[k][w] for w in want}, **kw)
            pend = []
            def run(n, i0):
                for i in range(n):
                    k = i % depth
                    if len(pend) >= depth:
                        pend.pop(0).result()
                    pend.append(pool.submit(lambda e, i=i, k=k: call(e, i0 + i, k), engine=k))
                while pend:
                    pend.pop(0).result()
            run(3, 0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(steps, 10)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / steps
            print("depth %d  %-42s %.2f ms/step  (%.2f M frames/s)" % (depth, name, ms, B * T / ms / 1e3), flush=True)
        pool.close()


if __name__ == "__main__":
    main()
