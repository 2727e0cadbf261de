"""Drop-in for `larynx.glow_tts.GlowTextToSpeech` (larynx/glow_tts.py:29-170) on libb200tts.

Same constructor (a TextToSpeechModelConfig whose model_path holds `config.json` + `generator.pth`, or `generator.onnx` whose
initializers kept their parameter names: pack.pack_from_directory),
same `phonemes_to_mels(phoneme_ids, settings)` contract: int64 ids [T_x] in, float32 mel
[1, 80, T_y] numpy array out (what the ONNX backend returns, larynx/glow_tts.py:161-170), same
settings keys (`noise_scale` 0.667, `length_scale` 1.0; larynx/glow_tts.py:104-121).
Extra, optional settings: `eps` (injected N(0,1) noise [80, >=T_y], for parity tests), `seed`, and `encoder_bf16`
(default: the instance's `encoder_bf16`, False) — the reduced-precision text encoder of BASELINE config 3.  `config.half`
is accepted and does NOT switch it on: it only acts on the reference's PyTorch-CUDA backend (larynx/glow_tts.py:90-91;
the ONNX backend this drop-in is measured against ignores it, :96-100) and larynx.load_tts_model defaults it to True
(larynx/__init__.py:298), so honouring it would silently drop every stock caller below the 1e-3 contract.
"""
from __future__ import annotations

import json
import logging
import os
import threading
import typing
from pathlib import Path

import numpy as np

from . import pack
from .constants import ARRAY_OR_TENSOR, SettingsType, TextToSpeechModel, TextToSpeechModelConfig
from .engine import CALLS, GlowModel, check_ids, fresh_seed

_LOGGER = logging.getLogger("larynx_b200.glow_tts")


def default_device() -> int:
    return int(os.environ.get("B200TTS_DEVICE", "0"))


class GlowTextToSpeech(TextToSpeechModel):
    def __init__(self, config: TextToSpeechModelConfig, device: typing.Optional[int] = None, blob=None,
                 model_config: typing.Optional[dict] = None):
        super().__init__(config)
        self.device = default_device() if device is None else int(device)
        model_path = Path(config.model_path) if config is not None else None
        if model_config is None:
            with open(model_path / "config.json", "r", encoding="utf-8") as f:
                model_config = json.load(f)
        self.config = model_config
        if blob is None:
            _LOGGER.debug("Packing GlowTTS weights from %s", model_path)
            blob = pack.pack_from_directory(model_path, "model", model_config, pack.pack_glow)
        self.model = GlowModel(blob, self.device)
        self.noise_scale = 0.667
        self.length_scale = 1.0
        self.encoder_bf16 = False  # opt-in: bf16 activations in the encoder GEMMs (alignments may differ; tools/half_report.py)
        self._seed_base = fresh_seed()  # per instance (process, rank, model): default noise never repeats across them
        self._seed_lock = threading.Lock()
        self._seed_next = 0

    # ---------------------------------------------------------------------------------------
    def phonemes_to_mels(self, phoneme_ids: np.ndarray, settings: typing.Optional[SettingsType] = None) -> ARRAY_OR_TENSOR:
        """Convert phoneme ids to mel spectrograms (float32 [1, mel, T_y])."""
        noise_scale, length_scale = self.noise_scale, self.length_scale
        eps = None
        seed = None
        bf16 = self.encoder_bf16
        if settings:
            bf16 = bool(settings.get("encoder_bf16", bf16))
            noise_scale = float(settings.get("noise_scale", noise_scale))
            length_scale = float(settings.get("length_scale", length_scale))
            if settings.get("speaker_id") is not None:
                raise ValueError("multi-speaker voices are not supported by larynx_b200")
            eps = settings.get("eps")
            seed = settings.get("seed")
        ids = np.ascontiguousarray(np.asarray(phoneme_ids, dtype=np.int64).reshape(1, -1))
        if ids.shape[1] == 0:
            return np.zeros((1, self.model.mel_channels, 0), dtype=np.float32)
        check_ids(ids, self.model.num_symbols)
        x_len = np.array([ids.shape[1]], dtype=np.int32)
        call = CALLS.get(self.device)
        y_len = self.model.encode(call, ids, x_len, 1, ids.shape[1], length_scale, encoder_bf16=bf16)
        t_y = int(y_len[0])
        mel = np.zeros((1, self.model.mel_channels, t_y), dtype=np.float32)
        if t_y == 0:
            return mel
        eps_ld = 0
        if eps is not None:
            eps = np.ascontiguousarray(np.asarray(eps, dtype=np.float32)[:, :t_y])
            assert eps.shape == (self.model.mel_channels, t_y), "eps must be [mel, >= T_y]"
            eps_ld = t_y
        self.model.decode(call, mel, t_y, noise_scale, eps, eps_ld, int(seed) if seed is not None else self.next_seed())
        return mel

    def next_seed(self) -> int:
        """Seed of the next default-noise call: instance base + call counter (explicit `seed` settings bypass it)."""
        with self._seed_lock:
            self._seed_next += 1
            return (self._seed_base + self._seed_next) & (2**64 - 1)
