"""runtime.AutoEngine (precision "auto"): tensor-core engine where it applies, fp32 engine otherwise - with the choice stated."""
import warnings

import numpy as np
import pytest

from oracle import reference_port as O
from tests.util import BF16_TOL, FP32_TOL, make_cfg, make_weights, max_abs, oracle_decode

pytestmark = pytest.mark.gpu


def test_default_configuration_takes_the_tensor_core_engine_and_falls_back_per_call():
    from gst_tacotron_b200.runtime import AutoEngine
    cfg = make_cfg("SMA")
    W = make_weights(cfg)
    eng = AutoEngine(cfg, W)
    try:
        assert eng.precision_used == "bf16"
        enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, 3, 40, 6, teacher=False)
        ref = oracle_decode(cfg, W, enc, steps=6, keep0=k0, keep1=k1, noise=nz)
        out = eng.decode(encodings=enc, steps=6, rng="external", keep0=k0, keep1=k1, noise=nz)
        assert max_abs(out["mel"], ref["decodings"]) < BF16_TOL and eng.fallback_calls == 0
        # 2000 keys do not fit the tensor-core kernel's shared memory: the call is repeated on the fp32 engine
        enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, 1, 2000, 3, teacher=False)
        ref = oracle_decode(cfg, W, enc, steps=3, keep0=k0, keep1=k1, noise=nz)
        out = eng.decode(encodings=enc, steps=3, rng="external", keep0=k0, keep1=k1, noise=nz)
        assert eng.fallback_calls == 1
        assert max_abs(out["mel"], ref["decodings"]) < FP32_TOL and max_abs(out["alignment"], ref["alignments"]) < FP32_TOL
        assert eng.launch_count > 0          # everything else is the wrapped engine's
    finally:
        eng.close()


def test_configuration_outside_the_tensor_core_path_gets_fp32_with_a_warning():
    from gst_tacotron_b200.runtime import AutoEngine
    cfg = make_cfg("SMA", lstm_sizes=[512, 512])
    W = make_weights(cfg)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        eng = AutoEngine(cfg, W)
    try:
        assert eng.precision_used == "fp32"
        assert any("tensor-core mode does not cover" in str(w.message) for w in rec)
        enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, 2, 21, 4, teacher=False)
        ref = oracle_decode(cfg, W, enc, steps=4, keep0=k0, keep1=k1, noise=nz)
        out = eng.decode(encodings=enc, steps=4, rng="external", keep0=k0, keep1=k1, noise=nz)
        assert max_abs(out["mel"], ref["decodings"]) < FP32_TOL
    finally:
        eng.close()
