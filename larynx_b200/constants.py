"""Plug-in boundary types.

When the reference package is importable (`larynx.constants`, larynx/constants.py:19-114) its own
classes are re-exported, so the drop-ins below are genuine subclasses of the reference ABCs.
Otherwise (e.g. on a box without the reference) structurally identical stand-ins are defined:
same names, same fields, same defaults.
"""
from __future__ import annotations

import typing
from abc import ABC
from dataclasses import field, make_dataclass
from enum import Enum
from pathlib import Path

import numpy as np

try:  # pragma: no cover - depends on the environment
    from larynx.constants import (  # type: ignore
        ARRAY_OR_TENSOR,
        InferenceBackend,
        SettingsType,
        TextToSpeechModel,
        TextToSpeechModelConfig,
        VocoderModel,
        VocoderModelConfig,
    )

    REFERENCE_TYPES = True
except Exception:  # noqa: BLE001 - larynx (or one of its imports) is not available
    REFERENCE_TYPES = False
    ARRAY_OR_TENSOR = typing.Union[np.ndarray, typing.Any]
    SettingsType = typing.Dict[str, typing.Any]

    InferenceBackend = Enum("InferenceBackend", {"ONNX": "onnx", "PYTORCH": "pytorch"}, type=str)

    # fields shared by both config records, in the reference's positional order (larynx/constants.py:51-59, 78-87)
    _COMMON = [("model_path", Path), ("session_options", typing.Any, field(default=None)),
               ("use_cuda", bool, field(default=True)), ("half", bool, field(default=True))]
    _BACKEND = ("backend", typing.Optional[InferenceBackend], field(default=None))
    TextToSpeechModelConfig = make_dataclass("TextToSpeechModelConfig", _COMMON + [_BACKEND])
    VocoderModelConfig = make_dataclass("VocoderModelConfig", _COMMON + [("denoiser_strength", float, field(default=0.0)), _BACKEND])

    class TextToSpeechModel(ABC):
        """phoneme ids -> mel spectrogram (larynx/constants.py:62-72)."""

        def __init__(self, config):
            self._config = config

        def phonemes_to_mels(self, phoneme_ids, settings=None):
            raise NotImplementedError

    class VocoderModel(ABC):
        """mel spectrogram -> int16 audio (larynx/constants.py:90-100)."""

        def __init__(self, config):
            self._config = config

        def mels_to_audio(self, mels, settings=None):
            raise NotImplementedError


# Generator architectures of the released vocoders (the `-q high | medium | low` choices of the reference CLI,
# hifi_gan/<name>/config.json): needed to build a vocoder from a bare generator.pth whose directory has no config.json.
HIFIGAN_CONFIGS = {
    "universal_large": {
        "resblock": "1",
        "upsample_rates": [8, 8, 2, 2],
        "upsample_kernel_sizes": [16, 16, 4, 4],
        "upsample_initial_channel": 512,
        "resblock_kernel_sizes": [3, 7, 11],
        "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    },
    "vctk_medium": {
        "resblock": "1",
        "upsample_rates": [8, 8, 2, 2],
        "upsample_kernel_sizes": [16, 16, 4, 4],
        "upsample_initial_channel": 128,
        "resblock_kernel_sizes": [3, 7, 11],
        "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    },
    "vctk_small": {
        "resblock": "2",
        "upsample_rates": [8, 8, 4],
        "upsample_kernel_sizes": [16, 16, 8],
        "upsample_initial_channel": 256,
        "resblock_kernel_sizes": [3, 5, 7],
        "resblock_dilation_sizes": [[1, 2], [2, 6], [3, 12]],
    },
}
