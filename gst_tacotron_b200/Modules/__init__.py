"""Drop-in mirrors of the reference's Keras layers on the decode hot path
(reference: Modules/Taco2.py, Modules/GST.py, Modules/Attention/{Steps,Layers}.py).

The reference's layers read ``Hyper_Parameters.json`` from the CWD at import and own Keras
variables that ``tf.train.Checkpoint`` restores.  Here the variables live in a weight pack
(``gst_tacotron_b200.weights``) inside a shared :class:`~gst_tacotron_b200.runtime.Engine`;
call :func:`configure` once (the equivalent of building the model + ``Restore()``), then use the
layer classes exactly like the reference's."""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np

from ..hparams import HotPathConfig, load_config
from ..runtime import Engine

_ENGINE: Optional[Engine] = None


def configure(weights: Mapping[str, np.ndarray], cfg: Optional[HotPathConfig] = None,
              hp_path: Optional[str] = None, device: int = 0, **overrides) -> Engine:
    """Create the process-wide engine used by layers constructed without ``engine=``."""
    global _ENGINE
    if cfg is None:
        cfg = load_config(hp_path, **overrides)
    _ENGINE = Engine(cfg, weights, device=device)
    return _ENGINE


def default_engine() -> Engine:
    if _ENGINE is None:
        raise RuntimeError("gst_tacotron_b200.Modules.configure(weights, ...) has not been called")
    return _ENGINE


def reset() -> None:
    global _ENGINE
    if _ENGINE is not None:
        _ENGINE.close()
    _ENGINE = None
