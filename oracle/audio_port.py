"""CPU ORACLE (test infrastructure, NOT product code) for the wav side of the reference: Audio.inv_spectrogram and the
Griffin-Lim loop it runs (Audio.py:23-27, 57-78, 102-106), called from Model.Export_Inference (Model.py:412-420).

The STFT / ISTFT themselves live in a third-party dependency that is absent here: ``librosa`` (the reference pins no
version; its API use - ``librosa.stft(y=, n_fft=, hop_length=, win_length=)``, ``librosa.istft(y, hop_length=, win_length=)``,
``np.complex`` - dates it to librosa 0.6-0.7 / numpy < 1.24).  Their published algorithm is restated below:
  stft   center=True: reflect-pad n_fft // 2 samples on both sides; frame t = padded[t * hop : t * hop + n_fft] * w with
         w = scipy.signal.get_window('hann', win_length, fftbins=True) (win_length == n_fft here); rfft -> [1 + n_fft/2, frames]
  istft  per frame irfft * w, overlap-add at t * hop into n_fft + hop * (frames - 1) samples, divide by the window sum of
         squares where it exceeds ``tiny``, drop n_fft // 2 samples at both ends
PARITY PINNING: tests/test_audio_cpu.py checks these two against torch.stft / torch.istft (center=True, pad_mode='reflect',
periodic Hann window), the one independent implementation of the same convention available in this image; librosa itself
could not be run, so everything downstream of it is pinned only that far.
"""
from __future__ import annotations

import numpy as np
from scipy import signal

MIN_LEVEL_DB = -100.0


def hann(n: int) -> np.ndarray:
    return signal.get_window("hann", n, fftbins=True)


def stft(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """librosa.stft(y, n_fft, hop_length=hop, win_length=n_fft): [1 + n_fft // 2, 1 + len(y) // hop] complex."""
    y = np.asarray(y, np.float64)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    w = hann(n_fft)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return np.fft.rfft(yp[idx] * w[None, :], axis=1).T


def istft(D: np.ndarray, hop: int, n_fft: int) -> np.ndarray:
    """librosa.istft(D, hop_length=hop, win_length=n_fft): hop * (frames - 1) samples."""
    n_frames = D.shape[1]
    w = hann(n_fft)
    frames = np.fft.irfft(D.T, n=n_fft, axis=1) * w[None, :]
    total = n_fft + hop * (n_frames - 1)
    y = np.zeros(total)
    wss = np.zeros(total)
    for t in range(n_frames):
        y[t * hop:t * hop + n_fft] += frames[t]
        wss[t * hop:t * hop + n_fft] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:total - n_fft // 2]


def denormalize(S, max_abs_value=None):
    """Audio._denormalize / _symmetric_denormalize (Audio.py:96-100)."""
    if max_abs_value is None:
        return np.clip(S, 0, 1) * -MIN_LEVEL_DB + MIN_LEVEL_DB
    return (np.clip(S, -max_abs_value, max_abs_value) + max_abs_value) / (2 * max_abs_value) * -MIN_LEVEL_DB + MIN_LEVEL_DB


def griffin_lim(S: np.ndarray, n_fft: int, hop: int, iters: int, init_uniform: np.ndarray) -> np.ndarray:
    """Audio._griffin_lim (Audio.py:57-68): S [1 + n_fft // 2, frames] magnitudes; ``init_uniform`` stands for the reference's
    np.random.rand(*S.shape) (:61), handed in so that the run is reproducible."""
    angles = np.exp(2j * np.pi * init_uniform)
    Sc = np.abs(S).astype(np.complex128)
    y = istft(Sc * angles, hop, n_fft)
    for _ in range(iters):
        angles = np.exp(1j * np.angle(stft(y, n_fft, hop)))
        y = istft(Sc * angles, hop, n_fft)
    return y


def inv_preemphasis(x, coef: float = 0.97):
    """Audio.inv_preemphasis (Audio.py:14-15)."""
    return signal.lfilter([1], [1, -coef], x)


def inv_spectrogram(spectrogram, num_freq, hop_length, win_length, sample_rate, ref_level_db=20, power=1.5, max_abs_value=None,
                    griffin_lim_iters=60, init_uniform=None):
    """Audio.inv_spectrogram (Audio.py:23-27): spectrogram [num_freq, frames] (the caller transposes the vocoder output,
    Model.py:413) -> waveform of hop * (frames - 1) samples."""
    S = denormalize(np.asarray(spectrogram), max_abs_value)   # in the caller's dtype, as numpy does for the reference (float32
    S = np.power(10.0, (S + ref_level_db) * 0.05)             # vocoder output: de-normalisation, amplitude and power in float32)
    n_fft = (num_freq - 1) * 2
    assert win_length == n_fft, "the reference always passes win_length == n_fft (Model.py:414-416)"
    return inv_preemphasis(griffin_lim(S ** power, n_fft, hop_length, griffin_lim_iters, init_uniform))
