"""Thin object layer over the C ABI: device-resident models + per-thread call contexts.

Nothing here computes: numpy/torch objects are only containers whose addresses are handed to
libb200tts.so (host or device pointers are both accepted by the library).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Sequence

import numpy as np

from . import _lib
from ._lib import Glue, check, ptr


class Call:
    """b200tts_call: one CUDA stream + workspaces.  Not thread-safe — use one per host thread."""

    def __init__(self, device: int = 0):
        self.device = int(device)
        self._h = C.c_void_p()
        check(_lib.load().b200tts_call_create(self.device, C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.load().b200tts_call_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def last_ms(self) -> float:
        v = C.c_float()
        check(_lib.load().b200tts_call_last_ms(self._h, C.byref(v)))
        return v.value

    @property
    def launches(self) -> int:
        v = C.c_int64()
        check(_lib.load().b200tts_call_launches(self._h, C.byref(v)))
        return v.value

    @property
    def stream(self) -> int:
        v = C.c_void_p()
        check(_lib.load().b200tts_call_stream(self._h, C.byref(v)))
        return v.value or 0

    def set_debug(self, on: bool):
        check(_lib.load().b200tts_call_set_debug(self._h, int(on)))

    def debug_noise(self, seed: int, batch: int, mel: int, n: int) -> np.ndarray:
        """The production Philox/Box-Muller stream on its own (what decode adds when no eps is injected)."""
        out = np.zeros((batch, mel, max(n, 1)), dtype=np.float32)
        check(_lib.load().b200tts_debug_noise(self._h, int(seed) & (2**64 - 1), batch, mel, n, ptr(out), out.shape[2]))
        return out[:, :, :n]

    def set_encoder_bf16(self, on: bool):
        """Reduced-precision text encoder for the following encode calls on this context (b200tts.h; default off)."""
        check(_lib.load().b200tts_call_set_encoder_bf16(self._h, int(bool(on))))

    def set_denoiser(self, strength: float):
        """Denoiser strength of the following vocoder calls on this context (0 = off; larynx/hifi_gan.py:152-166)."""
        check(_lib.load().b200tts_call_set_denoiser(self._h, float(strength)))


class _CallPool:
    """One Call per (host thread, device): `_sentence_task` runs on ThreadPoolExecutor workers against
    one shared model object (larynx/__init__.py:146-157), so the model must be re-entrant."""

    def __init__(self):
        self._tls = threading.local()

    def get(self, device: int) -> Call:
        d: Dict[int, Call] = getattr(self._tls, "calls", None)
        if d is None:
            d = {}
            self._tls.calls = d
        if device not in d:
            d[device] = Call(device)
        return d[device]


CALLS = _CallPool()


def _as_blob(blob):
    if isinstance(blob, np.ndarray):
        assert blob.dtype == np.uint8 and blob.flags["C_CONTIGUOUS"]
        return blob.ctypes.data, blob.nbytes
    if hasattr(blob, "data_ptr"):  # torch uint8 tensor, host or device (e.g. after an NCCL broadcast)
        assert blob.is_contiguous()
        return blob.data_ptr(), blob.numel() * blob.element_size()
    raise TypeError(type(blob))


class GlowModel:
    """b200tts_glow: GlowTTS weights resident on one device (immutable, shareable across threads)."""

    def __init__(self, blob, device: int = 0):
        self.device = int(device)
        self._h = C.c_void_p()
        addr, n = _as_blob(blob)
        check(_lib.load().b200tts_glow_create(addr, n, self.device, C.byref(self._h)))
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        check(_lib.load().b200tts_glow_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.num_symbols, self.mel_channels, self.n_sqz = a.value, b.value, c.value

    def close(self):
        if self._h:
            _lib.load().b200tts_glow_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def encode(self, call: Call, ids, x_len, batch: int, max_tx: int, length_scale: float = 1.0,
               encoder_bf16: bool = False) -> np.ndarray:
        """Phase A.  ids [batch, max_tx] int64, x_len [batch] int32 (numpy or device tensors)."""
        y_len = np.zeros(batch, dtype=np.int32)
        call.set_encoder_bf16(encoder_bf16)
        check(_lib.load().b200tts_glow_encode(self._h, call._h, ptr(ids), ptr(x_len), batch, max_tx, float(length_scale), ptr(y_len)))
        return y_len

    def decode(self, call: Call, mel_out, mel_ld: int, noise_scale: float = 0.667, eps=None, eps_ld: int = 0, seed: int = 0):
        check(_lib.load().b200tts_glow_decode(self._h, call._h, float(noise_scale), ptr(eps), int(eps_ld), int(seed), ptr(mel_out), int(mel_ld)))

    def alignment(self, call: Call, batch: int, max_tx: int, ty_max: int):
        logw = np.zeros((batch, max_tx), dtype=np.float32)
        w_ceil = np.zeros((batch, max_tx), dtype=np.int32)
        idx = np.full((batch, max(ty_max, 1)), -1, dtype=np.int32)
        check(_lib.load().b200tts_glow_alignment(self._h, call._h, ptr(logw), ptr(w_ceil), ptr(idx), idx.shape[1]))
        return logw, w_ceil, idx[:, :ty_max]

    def flow_tap(self, call: Call, blk: int, batch: int, tq: int) -> np.ndarray:
        out = np.zeros((batch, self.mel_channels * self.n_sqz, max(tq, 1)), dtype=np.float32)
        check(_lib.load().b200tts_glow_flow_tap(self._h, call._h, blk, ptr(out), out.shape[2]))
        return out[:, :, :tq]


class HifiGanModel:
    """b200tts_hifigan: HiFi-GAN generator weights resident on one device."""

    def __init__(self, blob, device: int = 0):
        self.device = int(device)
        self._h = C.c_void_p()
        addr, n = _as_blob(blob)
        check(_lib.load().b200tts_hifigan_create(addr, n, self.device, C.byref(self._h)))
        a, b = C.c_int32(), C.c_int32()
        check(_lib.load().b200tts_hifigan_info(self._h, C.byref(a), C.byref(b)))
        self.mel_channels, self.hop = a.value, b.value

    def close(self):
        if self._h:
            _lib.load().b200tts_hifigan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def bias_spectrum(self) -> np.ndarray:
        """Denoiser bias spectrum [513] (larynx/hifi_gan.py:181-203), computed on the device at model creation."""
        out = np.zeros(513, dtype=np.float32)
        check(_lib.load().b200tts_hifigan_bias_spectrum(self._h, ptr(out), 513))
        return out

    def denoise(self, call: Call, strength: float, pcm_in, n_samples, batch: int, in_ld: int, pcm_f32=None, pcm_i16=None,
                out_ld: int = 0, bias=None):
        check(_lib.load().b200tts_hifigan_denoise(self._h, call._h, float(strength), ptr(bias), ptr(pcm_in), ptr(n_samples), batch,
                                                  int(in_ld), ptr(pcm_f32), ptr(pcm_i16), int(out_ld)))

    def infer(self, call: Call, mel, frames, batch: int, mel_ld: int, pcm_f32=None, pcm_i16=None, pcm_ld: int = 0):
        check(_lib.load().b200tts_hifigan_infer(self._h, call._h, ptr(mel), ptr(frames), batch, int(mel_ld), ptr(pcm_f32), ptr(pcm_i16), int(pcm_ld)))


def tts_finish(glow: GlowModel, voc: HifiGanModel, call: Call, glue: Glue, noise_scale: float, eps, eps_ld: int, seed: int,
               mel_out, mel_ld: int, pcm_f32, pcm_i16, pcm_ld: int):
    check(_lib.load().b200tts_tts_finish(glow._h, voc._h, call._h, C.byref(glue), float(noise_scale), ptr(eps), int(eps_ld), int(seed),
                                         ptr(mel_out), int(mel_ld), ptr(pcm_f32), ptr(pcm_i16), int(pcm_ld)))


def fresh_seed() -> int:
    """A 64-bit seed nobody else has: the reference draws fresh torch.randn_like noise per call (glow_tts/models.py:348)."""
    return int.from_bytes(os.urandom(8), "little")


def check_ids(ids: np.ndarray, num_symbols: int):
    """nn.Embedding raises IndexError for ids outside the table (glow_tts/models.py:119); so do we, before any launch."""
    if ids.size and (int(ids.min()) < 0 or int(ids.max()) >= num_symbols):
        raise IndexError(f"phoneme id out of range for this voice: ids must be in [0, {num_symbols}), got "
                         f"[{int(ids.min())}, {int(ids.max())}]")


def pad_ids(id_rows: Sequence[np.ndarray]):
    """Ragged list of id arrays -> (ids [B, max_tx] int64 zero padded, x_len [B] int32)."""
    b = len(id_rows)
    max_tx = max((len(r) for r in id_rows), default=0)
    ids = np.zeros((b, max(max_tx, 1)), dtype=np.int64)
    x_len = np.zeros(b, dtype=np.int32)
    for i, r in enumerate(id_rows):
        r = np.asarray(r, dtype=np.int64).reshape(-1)
        ids[i, : len(r)] = r
        x_len[i] = len(r)
    return ids, x_len
