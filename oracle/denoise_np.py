"""TEST INFRASTRUCTURE — vectorised numpy restatement of the reference's spectral-subtraction denoiser.

Only tests/ may import this module (like everything under oracle/); the product path runs the denoiser on the
device (larynx_b200/csrc/kernels.cu, stft_denoise_kernel / overlap_add_kernel).  Pinned against the reference's
own `larynx.audio.transform` / `inverse` in tests/test_oracle_pinned.py and tests/golden/denoiser.npz.

Mirrors `HiFiGanVocoder.denoise` / `maybe_init_denoiser` (larynx/hifi_gan.py:171-203) and the STFT helpers
`transform` / `inverse` / `stft` / `istft` (larynx/audio.py:232-306): 1024-point periodic-less `np.hanning`
window, hop 256, frames `range(0, N - 1024, 256)`, magnitude minus `strength * bias_spec` clipped at 0, phase kept,
overlap-add of the windowed inverse FFT *without* window-power normalisation (the reference has none; the final
peak normalisation of `audio_float_to_int16` absorbs the gain), spectrum rounded through complex64 exactly where
the reference does (`larynx/audio.py:275-277`).

The reference runs this in per-frame Python loops on the host; here the same arithmetic is vectorised over frames
(oracle/port.py:denoise_reference keeps the loop form).
The bias spectrum is the first STFT frame of the vocoder's response to an all-zero mel of 88 frames.
"""
from __future__ import annotations

import numpy as np

FFT_SIZE = 1024
HOP = 256


def _frames(x: np.ndarray) -> np.ndarray:
    n = len(x)
    starts = np.arange(0, max(n - FFT_SIZE, 0), HOP)
    if len(starts) == 0:
        return np.zeros((0, FFT_SIZE), dtype=x.dtype)
    idx = starts[:, None] + np.arange(FFT_SIZE)[None, :]
    return x[idx]


def transform(audio: np.ndarray):
    """audio [1, N] -> (magnitude, phase), each [1, 513, n_frames] (larynx/audio.py:291-306)."""
    x = np.asarray(audio)[0]
    window = np.hanning(FFT_SIZE)
    spec = np.fft.rfft(window[None, :] * _frames(x), axis=1).T  # [513, n_t]
    real, imag = spec.real[None], spec.imag[None]
    return np.sqrt(real ** 2 + imag ** 2), np.arctan2(imag, real)


def inverse(magnitude: np.ndarray, phase: np.ndarray) -> np.ndarray:
    """(magnitude, phase) [1, 513, n_t] -> audio [1, n_t*256 + 1024] float64 (larynx/audio.py:270-288, 251-267)."""
    spec = np.empty(magnitude.shape, dtype=np.complex64)
    spec.real = magnitude * np.cos(phase)
    spec.imag = magnitude * np.sin(phase)
    n_t = spec.shape[2]
    window = np.hanning(FFT_SIZE)
    frames = window[None, :] * np.real(np.fft.irfft(spec[0].T, axis=1))  # [n_t, 1024]
    out = np.zeros(n_t * HOP + FFT_SIZE)
    # overlap-add: 4 interleaved groups of non-overlapping frames (hop = fft/4)
    for g in range(FFT_SIZE // HOP):
        sel = frames[g :: FFT_SIZE // HOP]
        if len(sel):
            start = g * HOP
            seg = out[start : start + len(sel) * FFT_SIZE].reshape(len(sel), FFT_SIZE)
            seg += sel
    return out[None, :]


def bias_spectrum(zero_mel_audio: np.ndarray) -> np.ndarray:
    """First STFT frame of the vocoder's response to zeros (larynx/hifi_gan.py:199-203): [1, 513, 1]."""
    mag, _ = transform(zero_mel_audio)
    return mag[:, :, 0][:, :, None]


def denoise(audio: np.ndarray, bias_spec: np.ndarray, strength: float) -> np.ndarray:
    """larynx/hifi_gan.py:171-179."""
    mag, phase = transform(audio)
    mag = np.clip(mag - bias_spec * strength, a_min=0.0, a_max=None)
    return inverse(mag, phase)


def audio_float_to_int16(audio: np.ndarray, max_wav_value: float = 32767.0) -> np.ndarray:
    """larynx/audio.py:118-125."""
    a = audio * (max_wav_value / max(0.01, np.max(np.abs(audio))))
    return np.clip(a, -max_wav_value, max_wav_value).astype("int16")
