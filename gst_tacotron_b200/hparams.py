"""Hyper_Parameters.json reader for the decode hot path.

The reference re-reads ``Hyper_Parameters.json`` from the CWD at import time in every
module (reference: Modules/Taco2.py:6-10, Modules/GST.py:6-10, Model.py:17-18).  This
module reads the *same file with the same key names* but only once, explicitly, and keeps
only the keys the hot path consumes (reference: Hyper_Parameters.json:4,13-38,109-132).

Optional keys that do not exist in the reference schema (all have defaults, so an
unmodified reference JSON loads unchanged):

``Tacotron2.Decoder.Attention.LSA``  {"Filters": 32, "Kernel_Size": 31, "Cumulate_Weights": true,
                                      "Smoothing": false}
    conv parameters for the step-form location-sensitive attention (the reference's
    LocationSensitiveAttention, Modules/Attention/Layers.py:289-320, takes them as
    constructor arguments and has no JSON entry for them).
``B200``  {"Precision": "fp32"|"bf16", "RNG": "external"|"philox", "Seed": int}
"""
from __future__ import annotations

import copy
import json
import os
from dataclasses import dataclass, field
from typing import List

# Shipped defaults == the values of the reference's Hyper_Parameters.json that the hot
# path reads (reference: Hyper_Parameters.json:2-38,93-132).
DEFAULT_HP = {
    "Sound": {"Spectrogram_Dim": 513, "Mel_Dim": 80, "Frame_Length": 1024, "Frame_Shift": 256,
              "Sample_Rate": 16000, "Max_Abs_Mel": 4},
    "GST": {
        "Use": True,
        "Reference_Encoder": {
            "Conv": {"Filters": [32, 32, 64, 64, 128, 128], "Kernel_Size": [3, 3, 3, 3, 3, 3],
                     "Strides": [2, 2, 2, 2, 2, 2]},
            "RNN": {"Size": 128},
            "Dense": {"Size": 128},
        },
        "Style_Token": {"Size": 16, "Embedding": {"Size": 256}, "Attention": {"Head": 4, "Size": 128}},
    },
    "Tacotron2": {
        "Encoder": {"Embedding": {"Size": 512},
                    "Conv": {"Filters": [512, 512, 512], "Kernel_Size": [5, 5, 5], "Strides": [1, 1, 1],
                             "Dropout_Rate": 0.5},
                    "RNN": {"Size": 256, "Zoneout": 0.0}},
        "Decoder": {
            "Prenet": {"Size": [256, 256], "Dropout_Rate": 0.5},
            "RNN": {"Size": [1024, 1024], "Zoneout": 0.0},
            "Attention": {"Type": "SMA", "Size": 128},
            "Conv": {"Filters": [512, 512, 512, 512], "Kernel_Size": [5, 5, 5, 5], "Strides": [1, 1, 1, 1],
                     "Dropout_Rate": 0.5},
        },
    },
    "Step_Reduction": 1,
    "Max_Step": 1000,
    "Use_Mixed_Precision": False,
}

ATTENTION_TYPES = ("SMA", "BMA", "LSA")


@dataclass
class HotPathConfig:
    """Flat view of the hyper-parameters the hot path needs."""

    mel_dim: int = 80
    step_reduction: int = 1
    max_step: int = 1000
    prenet_sizes: List[int] = field(default_factory=lambda: [256, 256])
    prenet_dropout: float = 0.5
    attention_type: str = "SMA"
    attention_size: int = 128
    lstm_sizes: List[int] = field(default_factory=lambda: [1024, 1024])
    zoneout: float = 0.0
    encoder_rnn_size: int = 256  # BiLSTM => 2x this many channels
    gst_use: bool = True
    ref_filters: List[int] = field(default_factory=lambda: [32, 32, 64, 64, 128, 128])
    ref_kernel: List[int] = field(default_factory=lambda: [3, 3, 3, 3, 3, 3])
    ref_strides: List[int] = field(default_factory=lambda: [2, 2, 2, 2, 2, 2])
    ref_gru_size: int = 128
    ref_dense_size: int = 128
    n_tokens: int = 16
    token_dim: int = 256
    style_heads: int = 4
    style_size: int = 128
    lsa_filters: int = 32
    lsa_kernel: int = 31
    lsa_cumulate: bool = True
    lsa_smoothing: bool = False
    # Postnet (Taco2.py:130-147): Conv.Filters + [Mel_Dim], Conv.Kernel_Size + [5], Conv.Strides + [1]
    postnet_filters: List[int] = field(default_factory=lambda: [512, 512, 512, 512])
    postnet_kernel: List[int] = field(default_factory=lambda: [5, 5, 5, 5])
    postnet_strides: List[int] = field(default_factory=lambda: [1, 1, 1, 1])
    # text Encoder (Taco2.py:12-51): Embedding -> Conv1D+BN+ReLU stack -> BiLSTM(encoder_rnn_size)
    vocab_size: int = 34          # len(Token_Index_Dict.ENG.json) (Taco2.py:9-10,19)
    encoder_embedding: int = 512
    encoder_filters: List[int] = field(default_factory=lambda: [512, 512, 512])
    encoder_kernel: List[int] = field(default_factory=lambda: [5, 5, 5])
    encoder_strides: List[int] = field(default_factory=lambda: [1, 1, 1])
    # Vocoder_Taco1 (Taco2.py:234-260, CBHG :285-385) + Griffin-Lim (Audio.py:23-27,57-68)
    spectrogram_dim: int = 513
    frame_length: int = 1024
    frame_shift: int = 256
    sample_rate: int = 16000
    max_abs_mel: float = 4.0      # Sound.Max_Abs_Mel: the max_abs_value Export_Inference hands to inv_spectrogram (Model.py:417)
    voc_bank_count: int = 8       # Conv1D kernel sizes 1 .. count
    voc_bank_filters: int = 256
    voc_pool_size: int = 2
    voc_pool_strides: int = 1
    voc_proj_filters: List[int] = field(default_factory=lambda: [128, 128])
    voc_proj_kernel: List[int] = field(default_factory=lambda: [3, 3])
    voc_highway_count: int = 4
    voc_highway_size: int = 128
    voc_rnn_size: int = 256
    griffin_lim_iters: int = 60
    precision: str = "fp32"
    rng: str = "external"
    seed: int = 0

    # ---- derived -------------------------------------------------------------------
    @property
    def text_dim(self) -> int:
        return 2 * self.encoder_rnn_size

    @property
    def gst_dim(self) -> int:
        return self.style_size if self.gst_use else 0

    @property
    def enc_dim(self) -> int:
        """Channel count of the decoder's `encodings` input (GST channels first,
        reference: Modules/GST.py:121-124)."""
        return self.text_dim + self.gst_dim

    @property
    def sigmoid_noise(self) -> float:
        # reference: Steps.py:58 (BMA default 0.0), Steps.py:212 (SMA default 2.0); LSA has none.
        return {"SMA": 2.0, "BMA": 0.0, "LSA": 0.0}[self.attention_type]

    @property
    def proj_dim(self) -> int:
        return self.mel_dim * self.step_reduction + 1  # reference: Taco2.py:87-89

    @property
    def postnet_layers(self):
        """[(filters, kernel_size, stride, tanh)] per Postnet layer, reference: Taco2.py:131-145."""
        f = list(self.postnet_filters) + [self.mel_dim]
        k = list(self.postnet_kernel) + [5]
        s = list(self.postnet_strides) + [1]
        n = len(self.postnet_filters)
        return [(f[i], k[i], s[i], i < n - 1) for i in range(min(len(f), len(k), len(s)))]

    @property
    def ref_compress(self) -> int:
        c = 1
        for s in self.ref_strides:
            c *= s
        return c  # reference: GST.py:38-40 (reduce_prod of strides)

    def validate(self) -> None:
        if self.attention_type not in ATTENTION_TYPES:
            # reference raises ValueError for anything but BMA/SMA (Taco2.py:75); LSA is the
            # north-star extension.
            raise ValueError("Unsupported attention type: {}".format(self.attention_type))
        if self.style_size % self.style_heads != 0:
            # reference: Layers.py:155-156
            raise ValueError("size must be divisible by num_heads. ('{}' % '{}' != 0)".format(
                self.style_size, self.style_heads))
        if self.zoneout != 0.0:
            raise ValueError("Zoneout != 0 is not supported on the inference hot path "
                             "(reference default 0.0, Hyper_Parameters.json:114-117)")


def _merge(base: dict, over: dict) -> dict:
    out = copy.deepcopy(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def load_hp_dict(path: str | None = None) -> dict:
    """Return the raw dict (reference-style ``hp_Dict``). ``path=None`` looks for
    ``Hyper_Parameters.json`` in the CWD like the reference does and falls back to the
    built-in defaults."""
    if path is None:
        path = "Hyper_Parameters.json" if os.path.exists("Hyper_Parameters.json") else None
    if path is None:
        return copy.deepcopy(DEFAULT_HP)
    with open(path, "r") as f:
        user = json.load(f)
    return _merge(DEFAULT_HP, user)


def _vocab_size(hp: dict) -> int:
    """len(token_Index_Dict) like Taco2.py:9-10,19 when Token_JSON_Path is readable from the CWD; 34 (the shipped ENG table) otherwise."""
    path = hp.get("Token_JSON_Path")
    if path and os.path.exists(path):
        with open(path, "r") as f:
            return len(json.load(f))
    return 34


def config_from_hp(hp: dict | None = None, **overrides) -> HotPathConfig:
    hp = _merge(DEFAULT_HP, hp or {})
    dec = hp["Tacotron2"]["Decoder"]
    gst = hp["GST"]
    lsa = dec["Attention"].get("LSA", {})
    b200 = hp.get("B200", {})
    voc = hp.get("Vocoder_Taco1", {})
    cb = voc.get("CBHG", {})
    precision = b200.get("Precision", "fp32")
    cfg = HotPathConfig(
        mel_dim=int(hp["Sound"]["Mel_Dim"]),
        step_reduction=int(hp["Step_Reduction"]),
        max_step=int(hp["Max_Step"]),
        prenet_sizes=[int(s) for s in dec["Prenet"]["Size"]],
        prenet_dropout=float(dec["Prenet"]["Dropout_Rate"]),
        attention_type=str(dec["Attention"]["Type"]),
        attention_size=int(dec["Attention"]["Size"]),
        lstm_sizes=[int(s) for s in dec["RNN"]["Size"]],
        zoneout=float(dec["RNN"].get("Zoneout", 0.0)),
        encoder_rnn_size=int(hp["Tacotron2"]["Encoder"]["RNN"]["Size"]),
        gst_use=bool(gst["Use"]),
        ref_filters=[int(v) for v in gst["Reference_Encoder"]["Conv"]["Filters"]],
        ref_kernel=[int(v) for v in gst["Reference_Encoder"]["Conv"]["Kernel_Size"]],
        ref_strides=[int(v) for v in gst["Reference_Encoder"]["Conv"]["Strides"]],
        ref_gru_size=int(gst["Reference_Encoder"]["RNN"]["Size"]),
        ref_dense_size=int(gst["Reference_Encoder"]["Dense"]["Size"]),
        n_tokens=int(gst["Style_Token"]["Size"]),
        token_dim=int(gst["Style_Token"]["Embedding"]["Size"]),
        style_heads=int(gst["Style_Token"]["Attention"]["Head"]),
        style_size=int(gst["Style_Token"]["Attention"]["Size"]),
        lsa_filters=int(lsa.get("Filters", 32)),
        lsa_kernel=int(lsa.get("Kernel_Size", 31)),
        lsa_cumulate=bool(lsa.get("Cumulate_Weights", True)),
        lsa_smoothing=bool(lsa.get("Smoothing", False)),
        postnet_filters=[int(v) for v in dec.get("Conv", {}).get("Filters", [512, 512, 512, 512])],
        postnet_kernel=[int(v) for v in dec.get("Conv", {}).get("Kernel_Size", [5, 5, 5, 5])],
        postnet_strides=[int(v) for v in dec.get("Conv", {}).get("Strides", [1, 1, 1, 1])],
        vocab_size=int(b200.get("Vocab_Size", _vocab_size(hp))),
        encoder_embedding=int(hp["Tacotron2"]["Encoder"]["Embedding"]["Size"]),
        encoder_filters=[int(v) for v in hp["Tacotron2"]["Encoder"]["Conv"]["Filters"]],
        encoder_kernel=[int(v) for v in hp["Tacotron2"]["Encoder"]["Conv"]["Kernel_Size"]],
        encoder_strides=[int(v) for v in hp["Tacotron2"]["Encoder"]["Conv"]["Strides"]],
        spectrogram_dim=int(hp["Sound"].get("Spectrogram_Dim", 513)),
        frame_length=int(hp["Sound"].get("Frame_Length", 1024)),
        frame_shift=int(hp["Sound"].get("Frame_Shift", 256)),
        sample_rate=int(hp["Sound"].get("Sample_Rate", 16000)),
        max_abs_mel=float(hp["Sound"].get("Max_Abs_Mel", 4)),
        voc_bank_count=int(cb.get("Conv_Bank", {}).get("Stack_Count", 8)),
        voc_bank_filters=int(cb.get("Conv_Bank", {}).get("Filters", 256)),
        voc_pool_size=int(cb.get("Pool", {}).get("Pool_Size", 2)),
        voc_pool_strides=int(cb.get("Pool", {}).get("Strides", 1)),
        voc_proj_filters=[int(v) for v in cb.get("Conv1D", {}).get("Filters", [128, 128])],
        voc_proj_kernel=[int(v) for v in cb.get("Conv1D", {}).get("Kernel_Size", [3, 3])],
        voc_highway_count=int(cb.get("Highwaynet", {}).get("Count", 4)),
        voc_highway_size=int(cb.get("Highwaynet", {}).get("Size", 128)),
        voc_rnn_size=int(cb.get("RNN", {}).get("Size", 256)),
        griffin_lim_iters=int(voc.get("Griffin-Lim_Iter", 60)),
        precision=str(precision),
        rng=str(b200.get("RNG", "external")),
        seed=int(b200.get("Seed", 0)),
    )
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError("unknown config field {}".format(k))
        setattr(cfg, k, v)
    cfg.validate()
    return cfg


def load_config(path: str | None = None, **overrides) -> HotPathConfig:
    return config_from_hp(load_hp_dict(path), **overrides)
