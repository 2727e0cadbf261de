"""Drop-in for `larynx.hifi_gan.HiFiGanVocoder` (larynx/hifi_gan.py:33-203) on libb200tts.

Same constructor `(config: VocoderModelConfig, executor=None)` and the same
`mels_to_audio(mels [1, 80, T] float32, settings) -> int16 [256*T]` contract
(larynx/hifi_gan.py:130-169), including the peak normalisation of
`audio_float_to_int16` (larynx/audio.py:118-125), which runs on the device.

The spectral-subtraction denoiser (larynx/hifi_gan.py:158-203, the CLI/server default strength is 0.005) runs on
the host in numpy exactly like the reference's (larynx_b200/denoise.py); its bias spectrum comes from one extra
vocoder call on an all-zero mel at first use.  A device-side STFT is SURVEY.md §8(f) "next".
"""
from __future__ import annotations

import json
import logging
import typing
from pathlib import Path

import numpy as np

from . import denoise as _denoise
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
            pytorch_path = model_path / "generator.pth"
            if not pytorch_path.is_file():
                raise FileNotFoundError(f"{pytorch_path} not found: larynx_b200 reads the PyTorch checkpoint of a vocoder")
            _LOGGER.debug("Packing HiFi-GAN weights from %s", pytorch_path)
            blob = pack.pack_hifigan(pack.load_state_dict(pytorch_path, "generator"), _effective_config(model_config))
        self.model = HifiGanModel(blob, self.device)
        self.mel_channels = self.model.mel_channels
        self.denoiser_strength = float(getattr(config, "denoiser_strength", 0.0) or 0.0) if config is not None else 0.0
        self.bias_spec: typing.Optional[np.ndarray] = None
        self._bias_lock = __import__("threading").Lock()

    def maybe_init_denoiser(self):
        """larynx/hifi_gan.py:181-203: bias spectrum = first STFT frame of vocoder(zeros[1, mel, 88])."""
        with self._bias_lock:
            if self.bias_spec is None:
                zeros = np.zeros((1, self.mel_channels, 88), dtype=np.float32)
                self.bias_spec = _denoise.bias_spectrum(self.mels_to_float(zeros)[None, :])

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
        if denoiser_strength > 0 and t > 0:
            # float waveform from the device, then the reference's host-side STFT denoiser and int16 conversion
            self.maybe_init_denoiser()
            audio = self.mels_to_float(m)[None, :]
            audio = _denoise.denoise(audio, self.bias_spec, denoiser_strength)
            return _denoise.audio_float_to_int16(audio).squeeze()
        audio = np.zeros(t * self.model.hop, dtype=np.int16)
        if t == 0:
            return audio
        frames = np.array([t], dtype=np.int32)
        self.model.infer(CALLS.get(self.device), m, frames, 1, t, None, audio, audio.shape[0])
        return audio

    def mels_to_float(self, mels: np.ndarray) -> np.ndarray:
        """Float waveform in (-1, 1) before int16 conversion (parity tests)."""
        m = np.ascontiguousarray(mels, dtype=np.float32).reshape(1, self.mel_channels, -1)
        t = m.shape[2]
        pcm = np.zeros(t * self.model.hop, dtype=np.float32)
        if t:
            self.model.infer(CALLS.get(self.device), m, np.array([t], dtype=np.int32), 1, t, pcm, None, pcm.shape[0])
        return pcm


def _effective_config(cfg: dict) -> dict:
    """The reference loads the vocoder config with dataclasses-json, which silently drops unknown keys:
    the flat jik876 json of universal_large (local/hifi_gan/universal_large/config.json) therefore yields
    the *defaults* of hifi_gan/config.py:29-40.  Nested configs ({"model": {...}}) are used as given."""
    from .synth import HIFIGAN_CONFIGS

    model = dict(HIFIGAN_CONFIGS["universal_large"])
    model.update(cfg.get("model", {}) if isinstance(cfg.get("model"), dict) else {})
    return {"model": model}
