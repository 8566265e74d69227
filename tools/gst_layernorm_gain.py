"""Why the style embedding is compared at 5e-4 while everything in front of it passes 1e-4: Layer_Norm (Layers.py:280-285) divides by
the standard deviation of its input row, so an absolute error e in front of it becomes e / std behind it.  Prints, for the fp32
engine on the test shapes: the error of the reference-encoder output and of the attention weights (both in front of the norm), the
smallest std of the normalised rows (oracle), and the error of the normalised output.   python tools/gst_layernorm_gain.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gst_tacotron_b200.runtime import Engine  # noqa: E402
from oracle import reference_port as O  # noqa: E402
from tests.util import make_cfg, make_weights, max_abs  # noqa: E402

cfg = make_cfg("SMA")
W = make_weights(cfg)
eng = Engine(cfg, W)
Wt = O.to_torch(W)
for B, T in ((1, 188), (4, 257), (3, 64), (5, 1000), (64, 500)):
    mels, lens = O.synth_gst_inputs(cfg, B, T, min_len=1)
    ref, ref_enc, ref_att = O.style_token_layer(W, cfg, mels, lens, return_parts=True)
    # the row Layer_Norm sees (oracle): attention result + projected query (Layers.py:211)
    q = O.dense(ref_enc[:, None, :], Wt[O.GST + "/Attention/Query/kernel"], Wt[O.GST + "/Attention/Query/bias"])
    tokens = torch.tanh(Wt[O.GST + "/gst_tokens"])[None].expand(B, -1, -1)
    v = O.dense(tokens, Wt[O.GST + "/Attention/Value/kernel"], Wt[O.GST + "/Attention/Value/bias"])
    H = cfg.style_heads
    qs, vs = torch.cat(torch.chunk(q, H, -1), 0), torch.cat(torch.chunk(v, H, -1), 0)
    res = torch.softmax(qs @ vs.transpose(1, 2), -1) @ vs
    pre = torch.cat(torch.chunk(res, H, 0), -1) + q
    std = pre.std(dim=-1, unbiased=False)
    out = eng.gst(mels, lens, drop_first=True, want=("gst", "ref", "attention"))
    e_ref, e_att, e_out = max_abs(out["ref"], ref_enc), max_abs(out["attention"], ref_att), max_abs(out["gst"], ref)
    print("B=%3d T=%4d: |err| reference encoder %.1e, attention %.1e | min row std %.3f -> gain <= %.1f | |err| after Layer_Norm %.1e" % (
        B, T, e_ref, e_att, float(std.min()), 1.0 / float(std.min()), e_out))
eng.close()
