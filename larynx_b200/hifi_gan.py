"""Drop-in for `larynx.hifi_gan.HiFiGanVocoder` (larynx/hifi_gan.py:33-203) on libb200tts.

Same constructor `(config: VocoderModelConfig, executor=None)` and the same
`mels_to_audio(mels [1, 80, T] float32, settings) -> int16 [256*T]` contract
(larynx/hifi_gan.py:130-169), including the peak normalisation of
`audio_float_to_int16` (larynx/audio.py:118-125), which runs on the device.

The spectral-subtraction denoiser (larynx/hifi_gan.py:158-203; the CLI/server default strength is 0.005) also runs
on the device: STFT, bias subtraction, inverse STFT and overlap-add are kernels of libb200tts that follow the vocoder
on the same stream (b200tts_call_set_denoiser), and the bias spectrum is computed when the model is created
(the reference computes it lazily on a worker thread, larynx/hifi_gan.py:116-128).
"""
from __future__ import annotations

import json
import logging
import typing
from pathlib import Path

import numpy as np

from . import pack
from .constants import ARRAY_OR_TENSOR, SettingsType, VocoderModel, VocoderModelConfig
from .engine import CALLS, HifiGanModel
from .glow_tts import default_device

_LOGGER = logging.getLogger("larynx_b200.hifi_gan")


class HiFiGanVocoder(VocoderModel):
    def __init__(self, config: VocoderModelConfig, executor=None, device: typing.Optional[int] = None, blob=None,
                 model_config: typing.Optional[dict] = None):
        super().__init__(config)
        self.device = default_device() if device is None else int(device)
        model_path = Path(config.model_path) if config is not None else None
        if model_config is None:
            with open(model_path / "config.json", "r", encoding="utf-8") as f:
                model_config = json.load(f)
        self.config = model_config
        if blob is None:
            _LOGGER.debug("Packing HiFi-GAN weights from %s", model_path)
            blob = pack.pack_from_directory(model_path, "generator", _effective_config(model_config), pack.pack_hifigan)
        self.model = HifiGanModel(blob, self.device)
        self.mel_channels = self.model.mel_channels
        self.denoiser_strength = float(getattr(config, "denoiser_strength", 0.0) or 0.0) if config is not None else 0.0
        self.bias_spec: typing.Optional[np.ndarray] = None

    def maybe_init_denoiser(self):
        """larynx/hifi_gan.py:181-203: bias spectrum [1, 513, 1] = first STFT frame of vocoder(zeros[1, mel, 88]).
        The device already holds it (computed at model creation); this only fetches a host copy."""
        if self.bias_spec is None:
            self.bias_spec = self.model.bias_spectrum().reshape(1, -1, 1)

    def denoise(self, audio: np.ndarray, denoiser_strength: float, bias_spec: typing.Optional[np.ndarray] = None) -> np.ndarray:
        """larynx/hifi_gan.py:171-179: audio [1, N] float -> denoised [1, n_t*256 + 1024] (float32 here; the reference's
        numpy code returns float64).  `bias_spec` [1, 513, 1] overrides the model's own bias spectrum."""
        a = np.ascontiguousarray(np.asarray(audio, dtype=np.float32).reshape(1, -1))
        n = a.shape[1]
        out = np.zeros((1, _denoised_len(n)), dtype=np.float32)
        bias = None if bias_spec is None else np.ascontiguousarray(np.asarray(bias_spec, dtype=np.float32).reshape(-1))
        assert bias is None or bias.shape[0] == 513
        self.model.denoise(CALLS.get(self.device), denoiser_strength, a, np.array([n], dtype=np.int32), 1, n, out, None,
                           out.shape[1], bias)
        return out

    def mels_to_audio(self, mels: ARRAY_OR_TENSOR, settings: typing.Optional[SettingsType] = None) -> np.ndarray:
        """Convert mel spectrograms to int16 audio."""
        if hasattr(mels, "detach"):
            mels = mels.detach().cpu().numpy()
        assert isinstance(mels, np.ndarray)
        denoiser_strength = self.denoiser_strength
        if settings:
            denoiser_strength = float(settings.get("denoiser_strength", denoiser_strength))
        m = np.ascontiguousarray(mels, dtype=np.float32)
        if m.ndim == 2:
            m = m[None]
        assert m.ndim == 3 and m.shape[0] == 1 and m.shape[1] == self.mel_channels, m.shape
        t = m.shape[2]
        denoise = denoiser_strength > 0 and t > 0
        audio = np.zeros(_denoised_len(t * self.model.hop) if denoise else t * self.model.hop, dtype=np.int16)
        if t == 0:
            return audio
        self._infer(m, t, None, audio, denoiser_strength if denoise else 0.0)
        return audio

    def mels_to_float(self, mels: np.ndarray, denoiser_strength: float = 0.0) -> np.ndarray:
        """Float waveform before int16 conversion (parity tests); with a strength > 0, after the denoiser."""
        m = np.ascontiguousarray(mels, dtype=np.float32).reshape(1, self.mel_channels, -1)
        t = m.shape[2]
        denoise = denoiser_strength > 0 and t > 0
        pcm = np.zeros(_denoised_len(t * self.model.hop) if denoise else t * self.model.hop, dtype=np.float32)
        if t:
            self._infer(m, t, pcm, None, denoiser_strength if denoise else 0.0)
        return pcm

    def _infer(self, m: np.ndarray, t: int, pcm_f32, pcm_i16, denoiser_strength: float):
        call = CALLS.get(self.device)
        frames = np.array([t], dtype=np.int32)
        n = t * self.model.hop
        out = pcm_f32 if pcm_f32 is not None else pcm_i16
        if out.shape[0] != n:
            # at most 4 frames with the denoiser: no STFT frame fits, the reference returns 1024 zeros
            # (larynx/audio.py:251-267 with zero frames) and `out` is already that
            return
        call.set_denoiser(denoiser_strength)
        try:
            self.model.infer(call, m, frames, 1, t, pcm_f32, pcm_i16, n)
        finally:
            call.set_denoiser(0.0)


def _denoised_len(n: int, fft: int = 1024, hop: int = 256) -> int:
    """Length of istft(stft(x)) for len(x) = n (larynx/audio.py:232-267): frames at 0, hop, ... < n - fft."""
    n_t = len(range(0, n - fft, hop))
    return n_t * hop + fft


def _effective_config(cfg: dict) -> dict:
    """The reference loads the vocoder config with dataclasses-json, which silently drops unknown keys:
    the flat jik876 json of universal_large (local/hifi_gan/universal_large/config.json) therefore yields
    the *defaults* of hifi_gan/config.py:29-40.  Nested configs ({"model": {...}}) are used as given."""
    from .constants import HIFIGAN_CONFIGS

    model = dict(HIFIGAN_CONFIGS["universal_large"])
    model.update(cfg.get("model", {}) if isinstance(cfg.get("model"), dict) else {})
    return {"model": model}
