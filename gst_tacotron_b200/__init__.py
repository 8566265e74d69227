"""gst_tacotron_b200 - B200-native (sm_100a) implementation of GST_Tacotron's inference hot path
(Tacotron2 decoder loop + GST front end) behind the reference's Keras-layer call signatures."""
from .hparams import HotPathConfig, load_config, config_from_hp, load_hp_dict  # noqa: F401
from .weights import init_weights, weight_spec, from_named_arrays, save_npz, load_npz  # noqa: F401

__all__ = ["HotPathConfig", "load_config", "config_from_hp", "load_hp_dict", "init_weights", "weight_spec",
           "from_named_arrays", "save_npz", "load_npz", "Engine"]


def __getattr__(name):
    if name == "Engine":  # lazy: importing the package must not require the CUDA library
        from .runtime import Engine
        return Engine
    raise AttributeError(name)
