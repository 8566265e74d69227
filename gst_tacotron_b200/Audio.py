"""``Audio.inv_spectrogram`` with the reference's signature (Audio.py:23-27), executed by libgsttaco.so: de-normalisation,
dB -> amplitude, ``** power``, Griffin-Lim (Audio.py:57-68) and the inverse pre-emphasis all run on the GPU
(csrc/griffin_lim.cuh).  ``inv_spectrograms`` is the batched form Export_Inference would use instead of its per-utterance loop
(Model.py:373-420).

The reference draws the initial phases from numpy's global generator (``np.random.rand``, Audio.py:61); here they come from the
library's counter-based Philox stream (``seed``), or from ``init_uniform`` when given - the same split as for the decoder's
dropout masks."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .Modules import default_engine
from .runtime import Engine


def inv_spectrogram(spectrogram, num_freq, hop_length, win_length, sample_rate, ref_level_db=20, power=1.5, max_abs_value=None,
                    griffin_lim_iters=60, seed: int = 0, init_uniform=None, engine: Optional[Engine] = None):
    """spectrogram [num_freq, frames] (normalised, as Model.py:413 hands it over) -> waveform [hop_length * (frames - 1)]."""
    spec = np.ascontiguousarray(np.asarray(spectrogram, np.float32).T[None])
    if spec.shape[2] != num_freq:
        raise ValueError("spectrogram must be [num_freq, frames]")
    uni = None if init_uniform is None else np.ascontiguousarray(np.asarray(init_uniform, np.float32).T[None])
    eng = engine or default_engine()
    wav = eng.griffin_lim(spec, iters=griffin_lim_iters, rng="philox" if uni is None else "external", seed=seed, init_uniform=uni,
                          ref_level_db=float(ref_level_db), power=float(power), max_abs_value=max_abs_value,
                          hop_length=hop_length, win_length=win_length, host_outputs=True)
    return wav[0]


def inv_spectrograms(spectrograms, lengths: Sequence[int], hop_length, win_length, ref_level_db=20, power=1.5, max_abs_value=None,
                     griffin_lim_iters=60, seed: int = 0, engine: Optional[Engine] = None):
    """Batched form: spectrograms [B, frames, num_freq] straight from Vocoder_Taco1, ``lengths[b]`` = max(1, stop index) *
    Step_Reduction (Model.py:380,413).  Returns the list of waveforms, utterance b with hop_length * (lengths[b] - 1) samples."""
    eng = engine or default_engine()
    wav = eng.griffin_lim(spectrograms, lengths=lengths, iters=griffin_lim_iters, rng="philox", seed=seed,
                          ref_level_db=float(ref_level_db), power=float(power), max_abs_value=max_abs_value,
                          hop_length=hop_length, win_length=win_length, host_outputs=True)
    return [wav[b, :max(0, hop_length * (int(n) - 1))] for b, n in enumerate(lengths)]
