"""Host-side mirror of the reference's per-sentence orchestration for the hot path.

`sentence_task` follows `larynx._sentence_task` (larynx/__init__.py:214-285): phonemes_to_mels ->
AudioSettings glue (numpy, host) -> mels_to_audio, with the same real-time-factor definition
(:266-268).  With the reference installed, `larynx.text_to_speech` calls the drop-in classes
directly; this module exists so the same flow can run — and be benchmarked —
where the reference package is absent, and to expose the fused on-device variant.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import engine
from ._lib import Glue


@dataclass
class MelGlue:
    """The three AudioSettings steps applied between the passes (larynx/audio.py:58-59,83-108),
    float32 numpy like the reference."""

    signal_norm: bool = False
    symmetric_norm: bool = True
    clip_norm: bool = True
    convert_db_to_amp: bool = True
    do_dynamic_range_compression: bool = True
    max_norm: float = 4.0
    min_level_db: float = -100.0
    ref_level_db: float = 20.0
    spec_gain: float = 1.0
    sample_rate: int = 22050

    @staticmethod
    def from_config(audio: dict) -> "MelGlue":
        keys = MelGlue.__dataclass_fields__.keys()
        return MelGlue(**{k: audio[k] for k in keys if k in audio})

    def apply(self, mels: np.ndarray) -> np.ndarray:
        m = np.asarray(mels, dtype=np.float32)
        if self.signal_norm:
            lo, hi = (-self.max_norm, self.max_norm) if self.symmetric_norm else (0.0, self.max_norm)
            if self.clip_norm:
                m = np.clip(m, lo, hi)
            if self.symmetric_norm:
                m = ((m + self.max_norm) * -self.min_level_db / (2 * self.max_norm)) + self.min_level_db
            else:
                m = (m * -self.min_level_db / self.max_norm) + self.min_level_db
            m = m + self.ref_level_db
        if self.convert_db_to_amp:
            m = np.power(10.0, m / self.spec_gain)
        if self.do_dynamic_range_compression:
            m = np.log(np.clip(m, a_min=1e-5, a_max=None))
        return m.astype(np.float32)

    def c_struct(self) -> Glue:
        return Glue.from_settings(self.__dict__)


@dataclass
class SentenceResult:
    audio: np.ndarray  # int16 [n]
    mel_frames: int
    infer_sec: float
    real_time_factor: float


def sentence_task(phoneme_ids, glue: MelGlue, tts_model, tts_settings, vocoder_model, vocoder_settings=None) -> SentenceResult:
    """Two boundary calls with the host glue in between — exactly the reference's flow."""
    t0 = time.perf_counter()
    mels = tts_model.phonemes_to_mels(phoneme_ids, settings=tts_settings)
    mels = glue.apply(mels)
    audio = vocoder_model.mels_to_audio(mels, settings=vocoder_settings)
    t1 = time.perf_counter()
    dur = audio.shape[-1] / glue.sample_rate
    infer = t1 - t0
    return SentenceResult(audio, mels.shape[-1], infer, infer / dur if dur > 0 else 0.0)


def synthesize_batch(tts_model, vocoder_model, glue: MelGlue, id_rows: Sequence[np.ndarray], noise_scale: float = 0.667,
                     length_scale: float = 1.0, eps: Optional[np.ndarray] = None, seed: Optional[int] = None, want_mel: bool = False,
                     want_float: bool = False, denoiser_strength: float = 0.0, call=None):
    """Fused path: ids -> PCM for a ragged batch without leaving the device between the passes
    (b200tts_glow_encode + b200tts_tts_finish).  Returns a list of int16 arrays (+ mels / float pcm).
    `denoiser_strength` > 0 runs the spectral-subtraction denoiser (larynx/hifi_gan.py:158-179) on the device
    after the vocoder, per item.  `seed` None draws a fresh noise stream (like the reference's torch.randn_like per call);
    `call` overrides the calling thread's own call context (the scheduler's workers pass theirs)."""
    g, v = tts_model.model, vocoder_model.model
    if call is None:
        call = engine.CALLS.get(g.device)
    if seed is None:
        seed = tts_model.next_seed() if hasattr(tts_model, "next_seed") else engine.fresh_seed()
    ids, x_len = engine.pad_ids(id_rows)
    engine.check_ids(ids, g.num_symbols)
    b = len(id_rows)
    y_len = g.encode(call, ids, x_len, b, ids.shape[1], length_scale, encoder_bf16=bool(getattr(tts_model, "encoder_bf16", False)))
    ty = int(y_len.max()) if b else 0
    n = ty * v.hop
    pcm_i16 = np.zeros((b, max(n, 1)), dtype=np.int16)
    pcm_f32 = np.zeros((b, max(n, 1)), dtype=np.float32) if want_float else None
    mel = np.zeros((b, g.mel_channels, max(ty, 1)), dtype=np.float32) if want_mel else None
    eps_ld = 0
    if eps is not None:
        eps = np.ascontiguousarray(eps, dtype=np.float32)
        assert eps.ndim == 3 and eps.shape[0] == b and eps.shape[2] >= ty
        eps_ld = eps.shape[2]
    call.set_denoiser(denoiser_strength)
    try:
        engine.tts_finish(g, v, call, glue.c_struct(), noise_scale, eps, eps_ld, int(seed) & (2**64 - 1), mel, mel.shape[2] if want_mel else 0,
                          pcm_f32, pcm_i16, pcm_i16.shape[1])
    finally:
        call.set_denoiser(0.0)

    def item(buf, i):
        n = int(y_len[i]) * v.hop
        if denoiser_strength > 0 and 0 < n <= 1024:  # no STFT frame fits: the reference returns 1024 zeros
            return np.zeros(1024, dtype=buf.dtype)
        return buf[i, :n]

    out = {"audio": [item(pcm_i16, i) for i in range(b)], "y_len": y_len}
    if want_mel:
        out["mel"] = [mel[i, :, : int(y_len[i])] for i in range(b)]
    if want_float:
        out["pcm"] = [item(pcm_f32, i) for i in range(b)]
    return out
