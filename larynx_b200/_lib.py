"""ctypes binding of libb200tts.so (include/b200tts.h).  Fails loudly — there is no fallback."""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path
from typing import Optional

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libb200tts.so"
_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


class B200TTSError(RuntimeError):
    pass


class Glue(C.Structure):
    """struct b200tts_glue — mirrors larynx.audio.AudioSettings fields used by larynx/__init__.py:242-249."""

    _fields_ = [
        ("signal_norm", C.c_int32),
        ("symmetric_norm", C.c_int32),
        ("clip_norm", C.c_int32),
        ("convert_db_to_amp", C.c_int32),
        ("do_dynamic_range_compression", C.c_int32),
        ("max_norm", C.c_float),
        ("min_level_db", C.c_float),
        ("ref_level_db", C.c_float),
        ("spec_gain", C.c_float),
    ]

    @staticmethod
    def from_settings(a) -> "Glue":
        get = (lambda k, d: a.get(k, d)) if isinstance(a, dict) else (lambda k, d: getattr(a, k, d))
        return Glue(
            int(bool(get("signal_norm", False))),
            int(bool(get("symmetric_norm", True))),
            int(bool(get("clip_norm", True))),
            int(bool(get("convert_db_to_amp", True))),
            int(bool(get("do_dynamic_range_compression", True))),
            float(get("max_norm", 4.0)),
            float(get("min_level_db", -100.0)),
            float(get("ref_level_db", 20.0)),
            float(get("spec_gain", 1.0)),
        )


_SIGNATURES = {
    "b200tts_last_error": (C.c_char_p, []),
    "b200tts_version": (C.c_int, []),
    "b200tts_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200tts_call_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "b200tts_call_destroy": (C.c_int, [C.c_void_p]),
    "b200tts_call_last_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "b200tts_call_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "b200tts_call_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200tts_call_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "b200tts_call_set_denoiser": (C.c_int, [C.c_void_p, C.c_float]),
    "b200tts_call_set_encoder_bf16": (C.c_int, [C.c_void_p, C.c_int]),
    "b200tts_glow_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "b200tts_glow_destroy": (C.c_int, [C.c_void_p]),
    "b200tts_glow_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b200tts_glow_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "b200tts_glow_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_int]),
    "b200tts_glow_alignment": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "b200tts_glow_flow_tap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "b200tts_debug_noise": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b200tts_debug_stage_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "b200tts_hifigan_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "b200tts_hifigan_destroy": (C.c_int, [C.c_void_p]),
    "b200tts_hifigan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b200tts_hifigan_infer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "b200tts_hifigan_denoise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int]),
    "b200tts_hifigan_bias_spectrum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200tts_tts_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Glue), C.c_float, C.c_void_p, C.c_int,
                                     C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path() -> Path:
    return Path(os.environ.get("B200TTS_LIB", str(_LIB_PATH)))


def load() -> C.CDLL:
    """dlopen libb200tts.so.  Raises if it has not been built — the product has no CPU path."""
    global _lib
    with _lock:
        if _lib is None:
            p = lib_path()
            if not p.exists():
                raise B200TTSError(
                    f"{p} not found: build it with `python -m larynx_b200.build` (nvcc, sm_100a). "
                    "larynx_b200 has no CPU fallback."
                )
            lib = C.CDLL(str(p))
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
        return _lib


def check(rc: int):
    if rc != 0:
        msg = load().b200tts_last_error()
        raise B200TTSError(msg.decode("utf-8", "replace") if msg else f"libb200tts call failed ({rc})")


def ptr(x) -> Optional[int]:
    """Raw address of a numpy array / torch tensor / int / None (host or device memory)."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    raise TypeError(f"cannot take the address of {type(x)}")


def device_count() -> int:
    n = C.c_int(0)
    check(load().b200tts_device_count(C.byref(n)))
    return n.value
