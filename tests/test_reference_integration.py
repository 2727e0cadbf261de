"""The drop-in boundary exercised by the REFERENCE'S OWN caller code (larynx/__init__.py:214-285, 379-407, 472-508).

`larynx.load_tts_model` / `larynx.load_vocoder_model` import `larynx.glow_tts` / `larynx.hifi_gan` lazily; INTEGRATION.md
shadows those two modules with `larynx_b200.glow_tts` / `larynx_b200.hifi_gan`.  This test does exactly that and then
runs the reference's unmodified factories and `_sentence_task` on four pool threads, against synthetic voice
directories in the reference's on-disk format (config.json + generator.pth), comparing with the same calls on the
reference's PyTorch backend.

Where it runs: only where /root/reference exists (the build container) — and there is no GPU there, so the device
layer below the drop-in classes (`engine.GlowModel` / `engine.HifiGanModel`, i.e. the ctypes -> C ABI calls) is replaced
by an oracle-backed stand-in (test infrastructure; allowed to import oracle/).  Everything ABOVE the C ABI is the product
code: module shadowing, constructors reading the voice directory, weight packing from generator.pth, settings
handling, id validation, shapes / dtypes, denoiser plumbing, re-entrancy from pool threads.  The C ABI itself is covered
by the -m gpu tests (test_parity_gpu.py, incl. the voice-directory constructors without the reference).
"""
import os
import sys
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

from larynx_b200 import synth
from oracle import port, ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="needs the reference at /root/reference")


class _FakeCall:
    def __init__(self, device=0):
        self.device = device
        self.denoise = 0.0
        self.last_ms = 0.0

    def set_denoiser(self, s):
        self.denoise = float(s)

    def set_debug(self, on):
        pass


class _FakeCalls:
    def __init__(self):
        self._tls = threading.local()
        self.threads = set()

    def get(self, device):
        if not hasattr(self._tls, "c"):
            self._tls.c = _FakeCall(device)
        self.threads.add(threading.get_ident())
        return self._tls.c


def _install_fakes(monkeypatch, loaded):
    """Replace the device layer under the drop-in classes with the oracle; record what the product packed."""
    from larynx_b200 import engine, glow_tts, hifi_gan, pack

    real_load = pack.load_state_dict

    def recording_load(path, key):
        sd = real_load(path, key)
        loaded[key] = sd
        return sd

    monkeypatch.setattr(pack, "load_state_dict", recording_load)
    calls = _FakeCalls()
    monkeypatch.setattr(glow_tts, "CALLS", calls)
    monkeypatch.setattr(hifi_gan, "CALLS", calls)

    class FakeGlow:
        def __init__(self, blob, device=0):
            assert isinstance(blob, np.ndarray) and blob.dtype == np.uint8 and blob[:8].tobytes() == pack.MAGIC
            self.device = device
            self.sd = port.to_torch_sd(loaded["model"])
            self.cfg = loaded["glow_cfg"]
            self.num_symbols = self.cfg["model"]["num_symbols"]
            self.mel_channels, self.n_sqz = 80, 2

        def _run(self, call, noise, eps):
            ids, xl, ls = call.enc
            return port.glow_forward(self.sd, self.cfg, torch.from_numpy(ids), torch.from_numpy(xl.astype(np.int64)), noise, ls,
                                     None if eps is None else torch.from_numpy(eps).unsqueeze(0))

        def encode(self, call, ids, x_len, batch, max_tx, length_scale=1.0, encoder_bf16=False):
            assert not encoder_bf16, "the stock factories must get the full-precision encoder (config.half is not honoured)"
            call.enc = (np.array(ids), np.array(x_len), float(length_scale))
            r = self._run(call, 0.0, None)
            return np.array([r["mel"].shape[2]], dtype=np.int32)

        def decode(self, call, mel_out, mel_ld, noise_scale=0.667, eps=None, eps_ld=0, seed=0):
            if eps is None:
                eps = np.random.RandomState(seed & 0xFFFFFFFF).randn(80, mel_ld).astype(np.float32)
            mel_out[...] = self._run(call, noise_scale, eps)["mel"].numpy()

    class FakeHifi:
        def __init__(self, blob, device=0):
            assert blob[:8].tobytes() == pack.MAGIC
            self.device = device
            self.sd = port.to_torch_sd(loaded["generator"])
            self.cfg = loaded["hifi_cfg"]
            self.mel_channels, self.hop = 80, 256

        def bias_spectrum(self):
            return port.denoiser_bias_spec(self.sd, self.cfg).reshape(-1).astype(np.float32)

        def infer(self, call, mel, frames, batch, mel_ld, pcm_f32=None, pcm_i16=None, pcm_ld=0):
            pcm = port.hifigan_forward(self.sd, self.cfg, torch.from_numpy(np.array(mel))).reshape(1, -1).numpy()
            if call.denoise > 0:
                pcm = port.denoise_reference(pcm, self.bias_spectrum().reshape(1, -1, 1).astype(np.float64), call.denoise)
            if pcm_f32 is not None:
                pcm_f32[...] = pcm.reshape(-1)
            if pcm_i16 is not None:
                pcm_i16[...] = port.float_to_int16(pcm).reshape(-1)

    monkeypatch.setattr(glow_tts, "GlowModel", FakeGlow)
    monkeypatch.setattr(hifi_gan, "HifiGanModel", FakeHifi)
    return calls


def test_reference_factories_and_sentence_task_run_on_the_drop_in(tmp_path, monkeypatch):
    ref_harness.import_reference()  # stubs + sys.path
    import larynx
    import larynx.audio
    from larynx.constants import InferenceBackend

    gdir = synth.write_glow_voice(tmp_path / "en-us" / "ljspeech-glow_tts")
    hdir = synth.write_hifigan_voice(tmp_path / "hifi_gan" / "universal_large")
    audio_settings = larynx.audio.AudioSettings(**synth.glow_config(46)["audio"])
    rows = [synth.bench_ids(n, seed=40 + n) for n in (9, 17, 5, 12)]
    eps = synth.seeded_eps(512, seed=3)
    tts_settings = {"noise_scale": 0.667, "length_scale": 1.0}

    def fake_randn_like(t, *a, **k):  # the reference's only RNG draw (glow_tts/models.py:348) -> the same eps as ours
        return torch.from_numpy(eps)[:, : t.shape[-1]].unsqueeze(0).to(t.dtype)

    # ---- the reference itself: PyTorch backend, its own classes
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_tts = larynx.load_tts_model("glow_tts", gdir, backend=InferenceBackend.PYTORCH, use_cuda=False, half=False)
        ref_voc = larynx.load_vocoder_model("hifi_gan", hdir, backend=InferenceBackend.PYTORCH, use_cuda=False, half=False,
                                            denoiser_strength=0.0)
    monkeypatch.setattr(torch, "randn_like", fake_randn_like)
    ref_audio = [larynx._sentence_task("t", ids, audio_settings, ref_tts, tts_settings, ref_voc, None) for ids in rows]
    monkeypatch.undo()

    # ---- the drop-in: shadow the two lazily imported modules, same factories, same caller
    import larynx_b200.glow_tts as g
    import larynx_b200.hifi_gan as h

    loaded = {"glow_cfg": synth.glow_config(46), "hifi_cfg": synth.hifigan_config("universal_large")}
    calls = _install_fakes(monkeypatch, loaded)
    monkeypatch.setitem(sys.modules, "larynx.glow_tts", g)
    monkeypatch.setitem(sys.modules, "larynx.hifi_gan", h)
    executor = ThreadPoolExecutor(4)
    tts = larynx.load_tts_model("glow_tts", gdir, backend=None, use_cuda=True, half=False)
    voc = larynx.load_vocoder_model("hifi_gan", hdir, denoiser_strength=0.005, executor=executor)
    assert type(tts).__module__ == "larynx_b200.glow_tts" and type(voc).__module__ == "larynx_b200.hifi_gan"
    assert set(loaded) >= {"model", "generator"}, "the constructors must read generator.pth of the voice directories"
    setattr(tts, "phoneme_to_id", {"a": 1})  # larynx/__init__.py:362-363
    setattr(tts, "audio_settings", audio_settings)
    assert voc.denoiser_strength == 0.005

    def task(ids):
        return larynx._sentence_task("t", ids, audio_settings, tts, dict(tts_settings, eps=eps), voc, {"denoiser_strength": 0.0},
                                     pause_before_ms=0, pause_after_ms=100)

    got = list(executor.map(task, rows))
    assert len(calls.threads) >= 2, "sentences must have run on several pool threads"
    pad = (100 * audio_settings.sample_rate) // 1000
    for a, r in zip(got, ref_audio):
        assert a.dtype == np.int16 and a.ndim == 1 and a.shape[0] == r.shape[0] + pad
        assert not a[-pad:].any()
        assert np.abs(a[:-pad].astype(np.int32) - r.astype(np.int32)).max() <= 2  # port vs reference: 1e-5 float
    # the vocoder's configured denoiser strength is the default of the settings key (larynx/hifi_gan.py:152-156)
    dn = larynx._sentence_task("t", rows[1], audio_settings, tts, dict(tts_settings, eps=eps), voc, None)
    assert dn.dtype == np.int16 and dn.shape[0] == (len(ref_audio[1]) - 1024 + 255) // 256 * 256 + 1024
    # errors surface as exceptions, like the reference's asserts
    with pytest.raises(IndexError):
        tts.phonemes_to_mels(np.array([1, 2, 46], dtype=np.int64))
    with pytest.raises(ValueError):
        tts.phonemes_to_mels(rows[0], {"speaker_id": 3})
    assert tts.phonemes_to_mels(np.zeros(0, dtype=np.int64)).shape == (1, 80, 0)
    executor.shutdown()
