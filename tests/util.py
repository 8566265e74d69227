"""Shared helpers for the parity tests (oracle = checker only)."""
import numpy as np
import torch

from gst_tacotron_b200.hparams import config_from_hp
from gst_tacotron_b200.weights import init_weights
from oracle import reference_port as O

FP32_TOL = 1e-4   # north_star: teacher-forced mel / alignment / stop within 1e-4 absolute in fp32
BF16_TOL = 1e-2   # north_star: 1e-2 in bf16


def make_cfg(attention="SMA", **over):
    hp = {"Tacotron2": {"Decoder": {"Attention": {"Type": attention, "Size": 128}}}}
    return config_from_hp(hp, **over)


def make_weights(cfg, seed=1234, bias_scale=0.05):
    return init_weights(cfg, seed=seed, bias_scale=bias_scale)


def oracle_decode(cfg, W, enc, mels=None, steps=None, keep0=None, keep1=None, noise=None, dtype=torch.float64):
    out = O.decoder_loop(W, cfg, enc, mels=mels, training=mels is not None, steps=steps,
                         keep0=keep0, keep1=keep1, noise=noise, dtype=dtype)
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def to_np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def max_abs(a, b):
    return float(np.max(np.abs(to_np(a).astype(np.float64) - to_np(b).astype(np.float64)))) if to_np(a).size else 0.0
