"""Plug-in boundary types.

When the reference package is importable (`larynx.constants`, larynx/constants.py:19-114) its own
classes are re-exported, so the drop-ins below are genuine subclasses of the reference ABCs.
Otherwise (e.g. on a box without the reference) structurally identical stand-ins are defined:
same names, same fields, same defaults.
"""
from __future__ import annotations

import typing
from abc import ABC
from dataclasses import dataclass
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

    class InferenceBackend(str, Enum):
        ONNX = "onnx"
        PYTORCH = "pytorch"

    @dataclass
    class TextToSpeechModelConfig:
        model_path: Path
        session_options: typing.Any = None
        use_cuda: bool = True
        half: bool = True
        backend: typing.Optional[InferenceBackend] = None

    class TextToSpeechModel(ABC):
        def __init__(self, config: TextToSpeechModelConfig):
            pass

        def phonemes_to_mels(self, phoneme_ids: np.ndarray, settings: typing.Optional[SettingsType] = None) -> ARRAY_OR_TENSOR:
            pass

    @dataclass
    class VocoderModelConfig:
        model_path: Path
        session_options: typing.Any = None
        use_cuda: bool = True
        half: bool = True
        denoiser_strength: float = 0.0
        backend: typing.Optional[InferenceBackend] = None

    class VocoderModel(ABC):
        def __init__(self, config: VocoderModelConfig):
            pass

        def mels_to_audio(self, mels: ARRAY_OR_TENSOR, settings: typing.Optional[SettingsType] = None) -> np.ndarray:
            pass
