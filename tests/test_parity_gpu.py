"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Contract (BASELINE.json north_star): mel and waveform within 1e-3 max-abs (fp32), alignment indices
bit-exact.  The fp32 kernels are far inside that; the tighter bounds below are what they achieve and
guard against regressions.  TOL_CONTRACT is asserted as well so the contract is written down.
"""
import os
import threading

import numpy as np
import pytest
import torch

from larynx_b200 import engine, host, synth
from oracle import port
from tests import common

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL_CONTRACT = 1e-3
TOL_MEL = 5e-5   # measured: 1e-5 (tensor-core bf16-split path), 1.5e-6 (fp32 SIMT path)
TOL_PCM = 2e-4   # measured: 6e-5 (tensor-core bf16-split path), 6e-6 (fp32 SIMT path)
TOL_E2E = 5e-4   # ids -> PCM through mel glue (which multiplies mel differences by 50 ln10): measured 1.2e-4


def _i16_tol(pcm_f32: np.ndarray, err: float = TOL_PCM) -> int:
    # the peak-normalisation gain 32767/max(0.01, peak) amplifies the float error
    return int(np.ceil(32767.0 * err / max(0.01, float(np.abs(pcm_f32).max())))) + 1


def _oracle(ids, noise=0.667, ls=1.0, eps=None):
    return port.sentence_task(common.oracle_glow_sd(), common.glow_cfg(), common.oracle_hifi_sd(), common.hifi_cfg(), ids, noise, ls, eps)


@pytest.mark.parametrize("name", ["ljspeech_fixture28.npz", "bench128.npz"])
def test_golden_vectors(name):
    g = np.load(os.path.join(GOLDEN, name))
    eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
    glow, voc = common.product_glow(), common.product_hifi()
    call = engine.CALLS.get(0)
    call.set_debug(True)
    mel = glow.phonemes_to_mels(g["ids"], {"noise_scale": float(g["noise_scale"]), "length_scale": float(g["length_scale"]), "eps": eps})
    call.set_debug(False)
    assert mel.dtype == np.float32 and mel.shape == (1, 80, g["mel"].shape[1])
    logw, w_ceil, idx = glow.model.alignment(call, 1, len(g["ids"]), mel.shape[2])
    assert np.array_equal(w_ceil[0], g["w_ceil"]), "durations must be bit-exact"
    assert np.array_equal(idx[0], g["idx"]), "alignment indices must be bit-exact"
    assert np.abs(logw[0] - g["logw"]).max() < 5e-5
    d = np.abs(mel[0] - g["mel"]).max()
    assert d <= TOL_MEL <= TOL_CONTRACT, d
    for blk in (11, 6, 0):
        if f"flow{blk}" in g:
            tap = glow.model.flow_tap(call, blk, 1, mel.shape[2] // 2)
            assert np.abs(tap[0] - g[f"flow{blk}"]).max() <= 5e-4  # |values| ~ 5: 1e-4 relative
    # boundary #2 on the golden glued mel
    pcm = voc.mels_to_float(g["mel_glued"][None])
    ref = g["pcm"]
    got = pcm if len(ref) == len(pcm) else pcm[::16]
    d = np.abs(got - ref).max()
    assert d <= TOL_PCM <= TOL_CONTRACT, d
    audio = voc.mels_to_audio(g["mel_glued"][None])
    assert audio.dtype == np.int16 and audio.shape == g["audio"].shape
    assert np.abs(audio.astype(np.int32) - g["audio"].astype(np.int32)).max() <= _i16_tol(pcm)


@pytest.mark.parametrize("name", ["ljspeech_fixture28.npz", "bench128.npz"])
def test_end_to_end_ids_to_pcm_contract(name):
    """The north-star contract on the COMPOSED path: phoneme ids -> mel -> glue -> PCM within 1e-3 max-abs of the reference's
    own chain (golden `pcm` = reference Generator on the reference's glued mel), through both public routes: the two
    boundary calls with the numpy glue in between (larynx._sentence_task's flow) and the fused on-device path."""
    g = np.load(os.path.join(GOLDEN, name))
    eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    ns, ls = float(g["noise_scale"]), float(g["length_scale"])
    mel = glow.phonemes_to_mels(g["ids"], {"noise_scale": ns, "length_scale": ls, "eps": eps})
    two_call = voc.mels_to_float(glue.apply(mel))
    fused = host.synthesize_batch(glow, voc, glue, [g["ids"]], ns, ls, eps=eps[None, :, : mel.shape[2]], want_float=True)
    ref = g["pcm"]
    sub = 1 if len(ref) == len(two_call) else 16  # bench128 stores every 16th sample of the reference waveform
    d2 = float(np.abs(two_call[::sub] - ref).max())
    df = float(np.abs(fused["pcm"][0][::sub] - ref).max())
    print(f"ids->pcm max|d|: two-call {d2:.3e}, fused {df:.3e} (contract {TOL_CONTRACT:g})")
    assert d2 <= TOL_E2E <= TOL_CONTRACT, d2
    assert df <= TOL_E2E <= TOL_CONTRACT, df
    audio = host.sentence_task(g["ids"], glue, glow, {"noise_scale": ns, "length_scale": ls, "eps": eps}, voc).audio
    assert audio.shape == g["audio"].shape
    assert np.abs(audio.astype(np.int32) - g["audio"].astype(np.int32)).max() <= _i16_tol(ref, TOL_E2E)


@pytest.mark.parametrize("n_ids,ls,noise", [(1, 1.0, 0.667), (2, 1.0, 0.0), (7, 1.5, 0.667), (64, 0.7, 0.3), (200, 1.0, 0.667), (300, 1.0, 0.667)])
def test_against_oracle(n_ids, ls, noise):
    ids = synth.bench_ids(n_ids, seed=100 + n_ids)
    eps = synth.seeded_eps(4096, seed=n_ids)
    o = _oracle(ids, noise, ls, eps)
    glow, voc = common.product_glow(), common.product_hifi()
    mel = glow.phonemes_to_mels(ids, {"noise_scale": noise, "length_scale": ls, "eps": eps})
    call = engine.CALLS.get(0)
    logw, w_ceil, idx = glow.model.alignment(call, 1, n_ids, mel.shape[2])
    w = np.exp(o["logw"].reshape(-1)) * ls
    if np.abs(w - np.round(w)).min() > 1e-4:  # otherwise a ceil() flip is legitimate fp32 noise
        assert np.array_equal(w_ceil[0], o["w_ceil"][0].astype(np.int32))
        assert np.array_equal(idx[0], o["idx"][0])
        assert mel.shape == o["mel"].shape
        assert np.abs(mel - o["mel"]).max() <= TOL_MEL
    pcm = voc.mels_to_float(o["mel_glued"])
    assert np.abs(pcm - o["pcm"]).max() <= TOL_PCM
    audio = voc.mels_to_audio(o["mel_glued"])
    assert np.abs(audio.astype(np.int32) - o["audio"].astype(np.int32)).max() <= _i16_tol(o["pcm"])


def test_edge_cases():
    glow, voc = common.product_glow(), common.product_hifi()
    assert glow.phonemes_to_mels(np.zeros(0, dtype=np.int64)).shape == (1, 80, 0)
    a = voc.mels_to_audio(np.zeros((1, 80, 0), dtype=np.float32))
    assert a.shape == (0,) and a.dtype == np.int16
    # a single frame through the vocoder (every conv is mostly zero padding)
    m1 = np.random.RandomState(0).randn(1, 80, 1).astype(np.float32)
    ref = port.hifigan_forward(common.oracle_hifi_sd(), common.hifi_cfg(), torch.from_numpy(m1)).reshape(-1).numpy()
    assert np.abs(voc.mels_to_float(m1) - ref).max() <= TOL_PCM
    # very short length scale: every duration rounds up to 1 frame, odd totals are truncated to even
    ids = synth.bench_ids(5, seed=3)
    o = _oracle(ids, 0.0, 0.05)
    mel = glow.phonemes_to_mels(ids, {"noise_scale": 0.0, "length_scale": 0.05})
    assert mel.shape == o["mel"].shape == (1, 80, 4)
    assert np.abs(mel - o["mel"]).max() <= TOL_MEL
    with pytest.raises(Exception):
        voc.model.infer(engine.CALLS.get(0), m1, np.array([5], dtype=np.int32), 1, 1, None, np.zeros(256, np.int16), 256)


@pytest.mark.parametrize("name", ["vctk_small", "vctk_medium"])
def test_vocoder_variants(name):
    g = np.load(os.path.join(GOLDEN, f"{name}_fixture28.npz"))
    voc = common.product_hifi(name)
    pcm = voc.mels_to_float(g["mel_glued"][None])
    assert pcm.shape == g["pcm"].shape
    assert np.abs(pcm - g["pcm"]).max() <= TOL_PCM


def test_other_symbol_table_voice_golden():
    """BASELINE config 4 serves voices whose `num_symbols` is not ljspeech's 46 (de 54, fr 42, ...): a reference-generated
    golden for a 54-symbol voice with ids beyond ljspeech's table, through the whole path."""
    g = np.load(os.path.join(GOLDEN, "de54_ids47.npz"))
    ns = int(g["num_symbols"])
    eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
    glow, voc = common.product_glow(ns, int(g["glow_seed"])), common.product_hifi()
    assert glow.model.num_symbols == 54
    call = engine.CALLS.get(0)
    mel = glow.phonemes_to_mels(g["ids"], {"noise_scale": float(g["noise_scale"]), "length_scale": float(g["length_scale"]), "eps": eps})
    logw, w_ceil, idx = glow.model.alignment(call, 1, len(g["ids"]), mel.shape[2])
    assert np.array_equal(w_ceil[0], g["w_ceil"]) and np.array_equal(idx[0], g["idx"])
    assert np.abs(logw[0] - g["logw"]).max() < 5e-5
    assert np.abs(mel[0] - g["mel"]).max() <= TOL_MEL
    glue = host.MelGlue.from_config(common.glow_cfg(ns)["audio"])
    pcm = voc.mels_to_float(glue.apply(mel))
    assert np.abs(pcm[::8] - g["pcm"]).max() <= TOL_E2E
    with pytest.raises(IndexError):  # the symbol table is per voice: 54 is out of range here, 53 is not
        glow.phonemes_to_mels(np.array([1, 54], dtype=np.int64))


@pytest.mark.parametrize("name", ["vctk_small", "vctk_medium"])
def test_vocoder_variants_at_bench_size_and_in_ragged_batches(name):
    g, b = np.load(os.path.join(GOLDEN, f"{name}_bench128.npz")), np.load(os.path.join(GOLDEN, "bench128.npz"))
    voc = common.product_hifi(name)
    pcm = voc.mels_to_float(b["mel_glued"][None])
    assert pcm.shape[0] == int(g["n_samples"])
    assert np.abs(pcm[::16] - g["pcm"]).max() <= TOL_PCM
    # ragged batch (ResBlock2 for vctk_small): every item equals its own per-utterance oracle run, no cross talk
    rs = np.random.RandomState(7)
    frames = np.array([33, 12, 50], dtype=np.int32)
    mel = (rs.randn(3, 80, 50) * 3 - 4).astype(np.float32)
    out = np.zeros((3, 50 * voc.model.hop), dtype=np.float32)
    voc.model.infer(engine.CALLS.get(0), mel, frames, 3, 50, out, None, out.shape[1])
    for i, t in enumerate(frames):
        ref = port.hifigan_forward(common.oracle_hifi_sd(name), common.hifi_cfg(name), torch.from_numpy(mel[i : i + 1, :, :t])).reshape(-1).numpy()
        assert np.abs(out[i, : t * voc.model.hop] - ref).max() <= TOL_PCM
        assert np.all(out[i, t * voc.model.hop :] == 0)


def test_ragged_batch_equals_per_utterance():
    glow, voc = common.product_glow(), common.product_hifi()
    rows = [synth.bench_ids(n, seed=s) for n, s in ((33, 1), (64, 2), (5, 3), (47, 4))]
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    eps1 = synth.seeded_eps(1024, seed=9)
    eps = np.ascontiguousarray(np.broadcast_to(eps1, (len(rows),) + eps1.shape))
    out = host.synthesize_batch(glow, voc, glue, rows, 0.667, 1.0, eps=eps, want_mel=True, want_float=True)
    for i, ids in enumerate(rows):
        o = _oracle(ids, 0.667, 1.0, eps1)
        assert out["mel"][i].shape == o["mel"][0].shape
        assert np.abs(out["mel"][i] - o["mel"][0]).max() <= TOL_MEL
        # per-item vocoder parity on the oracle's glued mel (the fused path glues its own mel: looser bound,
        # the glue multiplies mel differences by 50*ln10)
        assert np.abs(out["pcm"][i] - o["pcm"]).max() <= TOL_E2E
        single = host.synthesize_batch(glow, voc, glue, [ids], 0.667, 1.0, eps=eps[:1], want_float=True)
        assert np.array_equal(single["pcm"][0], out["pcm"][i]), "batched item must be bit-identical to a single call"
        assert np.array_equal(single["audio"][0], out["audio"][i])


def test_batch32_variable_lengths():
    """BASELINE config 3 shape: 32 utterances of 64..256 phonemes in one padded/ragged batch."""
    glow, voc = common.product_glow(), common.product_hifi()
    rs = np.random.RandomState(1234)
    lens = rs.randint(64, 257, size=32)
    rows = [synth.bench_ids(int(n), seed=300 + i) for i, n in enumerate(lens)]
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    eps1 = synth.seeded_eps(2048, seed=5)
    eps = np.ascontiguousarray(np.broadcast_to(eps1, (32,) + eps1.shape))
    out = host.synthesize_batch(glow, voc, glue, rows, 0.667, 1.0, eps=eps, want_mel=True, want_float=True)
    checked = 0
    for i in (0, 13, 31):
        o = _oracle(rows[i], 0.667, 1.0, eps1)
        w = np.exp(o["logw"].reshape(-1))
        if np.abs(w - np.round(w)).min() <= 1e-4:
            continue  # a legitimate ceil() flip is possible on this draw
        assert out["mel"][i].shape == o["mel"][0].shape
        assert np.abs(out["mel"][i] - o["mel"][0]).max() <= TOL_MEL
        assert np.abs(out["pcm"][i] - o["pcm"]).max() <= TOL_E2E
        checked += 1
    assert checked >= 2
    assert all(a.shape[0] == int(y) * 256 for a, y in zip(out["audio"], out["y_len"]))


TOL_DN = 3e-4  # denoised float waveform: vocoder error (<= TOL_PCM) through <= 1.5x overlap-add gain, fp32 FFT


def test_denoiser_stft_arithmetic_matches_reference_vector():
    """HiFiGanVocoder.denoise (larynx/hifi_gan.py:171-179) on the committed vector made by the reference's own
    larynx.audio.transform / inverse: random audio, a strong random bias (many bins clipped to zero)."""
    g = np.load(os.path.join(GOLDEN, "denoiser.npz"))
    voc = common.product_hifi()
    out = voc.denoise(g["audio"], float(g["strength"]), g["bias"])
    assert out.shape == g["out"].shape and out.dtype == np.float32
    assert np.abs(out - g["out"]).max() <= 2e-5 * max(1.0, float(np.abs(g["out"]).max()))
    # lengths that are not a multiple of the hop, shorter than one window, empty (larynx/audio.py:232-267)
    rs = np.random.RandomState(11)
    bias = np.abs(rs.randn(1, 513, 1)) * 0.05
    for n in (0, 300, 1024, 1025, 1281, 5000):
        a = (rs.randn(1, n) * 0.2).astype(np.float32)
        ref = port.denoise_reference(a, bias, 0.3)
        got = voc.denoise(a, 0.3, bias)
        assert got.shape == ref.shape, n
        assert np.abs(got - ref).max() <= 2e-5, n
    # strength 0 is STFT -> ISTFT: Hann applied twice, never normalised
    a = (rs.randn(1, 256 * 9) * 0.2).astype(np.float32)
    assert np.abs(voc.denoise(a, 0.0, bias) - port.denoise_reference(a, bias, 0.0)).max() <= 2e-5


def test_denoiser_matches_reference_flow():
    """CLI/server default: denoiser_strength = 0.005 (larynx/__main__.py:512-516) -> larynx/hifi_gan.py:158-179,
    against vectors made by the reference (vocoder on zeros -> bias spectrum; transform / clip / inverse; int16)."""
    g = np.load(os.path.join(GOLDEN, "ljspeech_fixture28.npz"))
    d = np.load(os.path.join(GOLDEN, "denoiser.npz"))
    voc = common.product_hifi()
    strength = float(d["fixture28_strength"])
    voc.maybe_init_denoiser()
    bias = d["fixture28_bias_spec"]
    assert voc.bias_spec.shape == bias.shape == (1, 513, 1)
    assert np.abs(voc.bias_spec - bias).max() <= 1e-4 * max(1.0, float(bias.max()))
    pcm = voc.mels_to_float(g["mel_glued"][None], strength)
    assert pcm.shape == d["fixture28_pcm_denoised"].shape
    assert np.abs(pcm - d["fixture28_pcm_denoised"]).max() <= TOL_DN
    audio = voc.mels_to_audio(g["mel_glued"][None], {"denoiser_strength": strength})
    ref = d["fixture28_audio_denoised"]
    assert audio.dtype == np.int16 and audio.shape == ref.shape
    assert np.abs(audio.astype(np.int32) - ref.astype(np.int32)).max() <= _i16_tol(d["fixture28_pcm_denoised"], TOL_DN)
    # the config attribute is the default of the settings key (larynx/hifi_gan.py:152-156)
    voc.denoiser_strength = strength
    try:
        assert np.array_equal(voc.mels_to_audio(g["mel_glued"][None]), audio)
    finally:
        voc.denoiser_strength = 0.0
    # and the next call on the same thread is not denoised
    plain = voc.mels_to_audio(g["mel_glued"][None])
    assert np.abs(plain.astype(np.int32) - g["audio"].astype(np.int32)).max() <= _i16_tol(g["pcm"])
    # at most 4 frames: no STFT frame fits, the reference returns 1024 zeros
    short = voc.mels_to_audio(g["mel_glued"][None, :, :3], {"denoiser_strength": strength})
    assert short.shape == (1024,) and not short.any()


def test_denoiser_in_fused_ragged_batch():
    """The fused ids -> PCM path with the denoiser on equals denoising each item's own waveform."""
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    rows = [synth.bench_ids(n, seed=300 + n) for n in (17, 1, 2, 40)]
    eps = np.stack([synth.seeded_eps(1024, seed=i) for i in range(len(rows))])
    plain = host.synthesize_batch(glow, voc, glue, rows, eps=eps, want_float=True)
    dn = host.synthesize_batch(glow, voc, glue, rows, eps=eps, want_float=True, denoiser_strength=0.01)
    bias = voc.model.bias_spectrum().reshape(1, -1, 1).astype(np.float64)
    for i in range(len(rows)):
        if plain["pcm"][i].shape[0] == 0:  # an item of zero frames stays empty (the reference cannot run it at all)
            assert dn["pcm"][i].shape == (0,) and dn["audio"][i].shape == (0,)
            continue
        ref = port.denoise_reference(plain["pcm"][i][None], bias, 0.01)
        assert dn["pcm"][i].shape == ref[0].shape, i
        assert np.abs(dn["pcm"][i] - ref[0]).max() <= 2e-5, i
        i16 = port.float_to_int16(dn["pcm"][i][None]).squeeze()
        assert np.abs(dn["audio"][i].astype(np.int32) - i16.astype(np.int32)).max() <= 1, i  # fp32 vs fp64 gain rounding


def test_vocoder_ragged_batch_has_no_cross_talk():
    voc = common.product_hifi()
    rs = np.random.RandomState(5)
    frames = np.array([40, 13, 27], dtype=np.int32)
    mel = (rs.randn(3, 80, 40) * 3 - 4).astype(np.float32)
    pcm = np.zeros((3, 40 * 256), dtype=np.float32)
    voc.model.infer(engine.CALLS.get(0), mel, frames, 3, 40, pcm, None, pcm.shape[1])
    for i, t in enumerate(frames):
        ref = port.hifigan_forward(common.oracle_hifi_sd(), common.hifi_cfg(), torch.from_numpy(mel[i : i + 1, :, :t])).reshape(-1).numpy()
        assert np.abs(pcm[i, : t * 256] - ref).max() <= TOL_PCM
        assert np.all(pcm[i, t * 256 :] == 0)


def test_fused_path_matches_two_call_path():
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    ids = common.fixture_ids("it_took_me_quite_a_long_time_to_develop_a_voice")
    eps = synth.seeded_eps(1024, seed=11)
    r = host.sentence_task(ids, glue, glow, {"noise_scale": 0.667, "eps": eps}, voc)
    f = host.synthesize_batch(glow, voc, glue, [ids], 0.667, 1.0, eps=eps[None], want_float=True)
    o = _oracle(ids, 0.667, 1.0, eps)
    assert r.audio.shape == f["audio"][0].shape == o["audio"].shape
    tol = _i16_tol(o["pcm"], TOL_E2E)
    assert np.abs(r.audio.astype(np.int32) - o["audio"].astype(np.int32)).max() <= tol
    assert np.abs(f["audio"][0].astype(np.int32) - o["audio"].astype(np.int32)).max() <= tol
    assert r.real_time_factor > 0


def test_full_size_properties():
    """Size-independent properties at the benchmark size (128 ids / ~500 frames)."""
    glow, voc = common.product_glow(), common.product_hifi()
    ids = synth.bench_ids(128)
    eps = synth.seeded_eps(1024)
    s = {"noise_scale": 0.667, "eps": eps}
    m1 = glow.phonemes_to_mels(ids, s)
    m2 = glow.phonemes_to_mels(ids, s)
    assert np.array_equal(m1, m2), "deterministic: identical inputs give bit-identical mels"
    glued = host.MelGlue.from_config(common.glow_cfg()["audio"]).apply(m1)
    full = voc.mels_to_float(glued)
    assert np.array_equal(full, voc.mels_to_float(glued))
    assert np.all(np.isfinite(full)) and np.abs(full).max() <= 1.0
    # locality: the vocoder's receptive field is ~12 frames per side — the interior of
    # a cropped utterance must reproduce the same samples
    a, b = 100, 300
    crop = voc.mels_to_float(glued[:, :, a:b])
    lo, hi = 24 * 256, (b - a - 24) * 256
    assert np.abs(crop[lo:hi] - full[a * 256 + lo : a * 256 + hi]).max() <= 1e-5
    # durations scale: total frames is monotone in length_scale
    t1 = glow.phonemes_to_mels(ids, {"noise_scale": 0.0, "length_scale": 0.8}).shape[2]
    t2 = glow.phonemes_to_mels(ids, {"noise_scale": 0.0, "length_scale": 1.25}).shape[2]
    assert t1 < m1.shape[2] < t2


def test_philox_noise_stream_moments_and_wiring():
    """The production noise (no injected eps; glow_tts/models.py:348 draws torch.randn_like): the Philox4x32-10 + Box-Muller
    stream must be N(0,1) — mean, variance, 4th moment, tails, no serial / cross-channel / cross-seed correlation — and decode
    must add exactly that stream."""
    glow = common.product_glow()
    call = engine.CALLS.get(0)
    x = call.debug_noise(seed=0x1234567, batch=2, mel=80, n=8192).astype(np.float64)  # 1.3e6 draws
    n = x.size
    se = 1.0 / np.sqrt(n)
    assert np.isfinite(x).all()
    assert abs(x.mean()) < 5 * se
    assert abs(x.var() - 1.0) < 0.01
    assert abs((x ** 4).mean() - 3.0) < 0.05            # kurtosis of a normal; sd of the estimate ~ 0.0086
    assert abs((x ** 3).mean()) < 0.02                   # skewness; sd ~ 0.0034
    assert abs(np.mean(np.abs(x) > 3.0) - 0.0026998) < 0.0003
    assert np.abs(x).max() < 6.5                         # 24-bit uniforms bound the tail at sqrt(-2 ln 2^-25) = 5.9
    lag1 = np.mean(x[:, :, 1:] * x[:, :, :-1])
    chan = np.mean(x[:, 1:, :] * x[:, :-1, :])
    item = np.mean(x[0] * x[1])
    assert abs(lag1) < 5 * se and abs(chan) < 5 * se and abs(item) < 5 * se * np.sqrt(2)
    y = call.debug_noise(seed=0x1234568, batch=2, mel=80, n=8192).astype(np.float64)
    assert abs(np.mean(x * y)) < 5 * se
    # per-channel means / variances: no channel is special
    assert np.abs(x.mean(axis=(0, 2))).max() < 6 / np.sqrt(2 * 8192)
    assert np.abs(x.var(axis=(0, 2)) - 1.0).max() < 0.08
    # wiring: decode without eps == decode with this stream injected, bit for bit; seeds select streams
    ids = synth.bench_ids(64, seed=5)
    a = glow.phonemes_to_mels(ids, {"noise_scale": 0.667, "seed": 42})
    eps = call.debug_noise(seed=42, batch=1, mel=80, n=a.shape[2])[0]
    b = glow.phonemes_to_mels(ids, {"noise_scale": 0.667, "eps": eps})
    assert np.array_equal(a, b)
    c = glow.phonemes_to_mels(ids, {"noise_scale": 0.667, "seed": 43})
    base = glow.phonemes_to_mels(ids, {"noise_scale": 0.0})
    assert not np.array_equal(a, c) and not np.array_equal(a, base)
    # default seeds: two calls draw different noise (the reference draws fresh noise per call)
    d1, d2 = glow.phonemes_to_mels(ids), glow.phonemes_to_mels(ids)
    assert d1.shape == d2.shape and not np.array_equal(d1, d2)


def test_out_of_range_ids_raise():
    """nn.Embedding raises IndexError on ids outside the table (glow_tts/models.py:119); a voice / phoneme-map mismatch must
    not produce plausible audio silently — neither for host ids (checked before launch) nor for device-resident ids."""
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    bad = np.array([3, 4, 46, 5], dtype=np.int64)
    with pytest.raises(IndexError):
        glow.phonemes_to_mels(bad)
    with pytest.raises(IndexError):
        host.synthesize_batch(glow, voc, glue, [np.array([1, 2, 3]), np.array([-1, 2])])
    call = engine.CALLS.get(0)
    ids_d = torch.tensor([[3, 4, 99, 5]], dtype=torch.int64, device="cuda:0")
    xlen_d = torch.tensor([4], dtype=torch.int32, device="cuda:0")
    with pytest.raises(Exception, match="out of range"):
        glow.model.encode(call, ids_d, xlen_d, 1, 4, 1.0)
    # the failed encode left no state behind: decode refuses, the next good call works
    with pytest.raises(Exception):
        glow.model.decode(call, np.zeros((1, 80, 8), np.float32), 8, 0.0, None, 0, 1)
    assert glow.phonemes_to_mels(np.array([3, 4, 45, 5], dtype=np.int64), {"noise_scale": 0.0}).shape[1] == 80


def test_voice_directory_constructors_on_pool_threads(tmp_path):
    """The reference's constructor path (larynx/__init__.py:379-407, 472-508): a config object whose model_path holds
    config.json + generator.pth, `executor=` for the vocoder, `denoiser_strength` on the config — then `_sentence_task`'s
    flow from four pool threads.  (tests/test_reference_integration.py drives the same classes through the reference's
    own factories where the reference exists.)"""
    from concurrent.futures import ThreadPoolExecutor

    from larynx_b200.constants import TextToSpeechModelConfig, VocoderModelConfig
    from larynx_b200.glow_tts import GlowTextToSpeech
    from larynx_b200.hifi_gan import HiFiGanVocoder

    gdir = synth.write_glow_voice(tmp_path / "en-us" / "ljspeech-glow_tts")
    hdir = synth.write_hifigan_voice(tmp_path / "hifi_gan" / "universal_large")
    pool = ThreadPoolExecutor(4)
    tts = GlowTextToSpeech(TextToSpeechModelConfig(model_path=gdir, session_options=None, use_cuda=True, half=False))
    voc = HiFiGanVocoder(VocoderModelConfig(model_path=hdir, session_options=None, denoiser_strength=0.005), executor=pool)
    setattr(tts, "phoneme_to_id", {"a": 1})  # larynx/__init__.py:362-363
    assert voc.denoiser_strength == 0.005
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    eps = synth.seeded_eps(512, seed=3)
    rows = [synth.bench_ids(20 + 5 * i, seed=60 + i) for i in range(8)]
    # same weights, same seeds as the blob-constructed fixtures: identical results
    blob_tts, blob_voc = common.product_glow(), common.product_hifi()
    want = [host.sentence_task(r, glue, blob_tts, {"eps": eps}, blob_voc, {"denoiser_strength": 0.005}).audio for r in rows]
    got = list(pool.map(lambda r: host.sentence_task(r, glue, tts, {"eps": eps}, voc).audio, rows))
    for w, g_ in zip(want, got):
        assert g_.dtype == np.int16 and np.array_equal(w, g_)
    pool.shutdown()


def test_reentrant_from_host_threads():
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    rows = [synth.bench_ids(20 + 3 * i, seed=50 + i) for i in range(6)]
    eps = synth.seeded_eps(512, seed=3)
    expect = [host.sentence_task(r, glue, glow, {"eps": eps}, voc).audio for r in rows]
    got = [None] * len(rows)

    def work(i):
        for _ in range(3):
            got[i] = host.sentence_task(rows[i], glue, glow, {"eps": eps}, voc).audio

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(rows))]
    [t.start() for t in th]
    [t.join() for t in th]
    for e, g_ in zip(expect, got):
        assert np.array_equal(e, g_)


@pytest.mark.timeout(300)
def test_concurrent_batches_from_two_threads_finish():
    """Two host threads, each with its own call (stream), run micro-batches of long sentences at the same time: the persistent
    kernels of both (encoder chain, flow chain, one dataflow launch per HiFi-GAN stage) need every CTA resident; without the
    device-wide gate (kernels.cu: spin_gate) two such grids could interleave and wait for each other forever (bench.py
    config 5 hung).  Results must equal the single-threaded ones."""
    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    rows = [[synth.bench_ids(int(n), seed=900 + 17 * t + i) for i, n in enumerate((96, 120, 77, 128, 64, 101, 88, 115))] for t in range(2)]
    expect = [host.synthesize_batch(glow, voc, glue, r, seed=11)["audio"] for r in rows]
    got = [None, None]

    def work(t):
        for _ in range(12):
            got[t] = host.synthesize_batch(glow, voc, glue, rows[t], seed=11)["audio"]

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for e, g_ in zip(expect, got):
        assert len(e) == len(g_)
        for a, b_ in zip(e, g_):
            assert np.array_equal(a, b_)


def test_simt_reference_path_still_matches():
    """The fp32 SIMT conv path (B200TTS_DEBUG=simt) is kept as an in-library cross-check of the tensor-core path."""
    import subprocess
    import sys

    code = (
        "import sys, numpy as np; sys.path.insert(0, '.');"
        "from tests import common; from larynx_b200 import synth, host;"
        "ids = synth.bench_ids(40, seed=21); eps = synth.seeded_eps(512, seed=4);"
        "g, v = common.product_glow(), common.product_hifi();"
        "glue = host.MelGlue.from_config(common.glow_cfg()['audio']);"
        "o = host.synthesize_batch(g, v, glue, [ids], eps=eps[None], want_mel=True, want_float=True);"
        "np.savez(sys.argv[1], mel=o['mel'][0], pcm=o['pcm'][0])"
    )
    import os
    import tempfile

    outs = {}
    for mode in ("simt", "umma"):
        f = os.path.join(tempfile.mkdtemp(), mode + ".npz")
        env = dict(os.environ, B200TTS_DEBUG="simt" if mode == "simt" else "")
        subprocess.run([sys.executable, "-c", code, f], check=True, env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
        outs[mode] = np.load(f)
    assert outs["simt"]["mel"].shape == outs["umma"]["mel"].shape
    assert np.abs(outs["simt"]["mel"] - outs["umma"]["mel"]).max() <= TOL_MEL
    assert np.abs(outs["simt"]["pcm"] - outs["umma"]["pcm"]).max() <= TOL_E2E  # each path glues its own mel


def test_vocoder_launch_schemes_agree_bit_for_bit():
    """The HiFi-GAN ResBlocks run as one dataflow launch per stage (stage_umma.cuh, the default; the 32-channel stage with a
    whole ResBlock step per item unless no_stage_pair), as fused c1->c2 steps for the narrow stages (rbpair_umma.cuh,
    B200TTS_DEBUG=no_stage) or as one launch per conv position (no_stage,no_rbpair), and the stage kernels cut their tiles into
    per-CTA ranges in striped or (no_stage_stripe) contiguous order.  All of them
    execute the same MMAs in the same order with the same epilogue arithmetic: the waveforms must be identical, for a single
    utterance and for a ragged batch (different tile ranges per item, tiles that end inside a 128-row block)."""
    import subprocess
    import sys
    import tempfile

    code = """
import sys
import numpy as np
sys.path.insert(0, '.')
from tests import common
from larynx_b200 import engine
v = common.product_hifi()
call = engine.CALLS.get(0)
rs = np.random.RandomState(5)
lens = [131, 47, 300]
T = max(lens)
batch = np.zeros((3, 80, T), np.float32)
for i, n in enumerate(lens):
    batch[i, :, :n] = rs.randn(80, n) * 2 - 5
one = v.mels_to_float(batch[0:1, :, :lens[0]])
hop = v.model.hop
many = np.zeros((3, T * hop), np.float32)
v.model.infer(call, batch, np.array(lens, np.int32), 3, T, many, None, T * hop)
np.savez(sys.argv[1], one=one, many=many)
"""
    outs = {}
    for mode in ("", "no_stage_stripe", "no_stage_pair", "no_stage", "no_stage,no_rbpair"):
        f = os.path.join(tempfile.mkdtemp(), "o.npz")
        subprocess.run([sys.executable, "-c", code, f], check=True, env=dict(os.environ, B200TTS_DEBUG=mode),
                       cwd=os.path.dirname(os.path.dirname(__file__)))
        outs[mode] = np.load(f)
    for mode in ("no_stage_stripe", "no_stage_pair", "no_stage", "no_stage,no_rbpair"):
        assert np.array_equal(outs[""]["one"], outs[mode]["one"]), mode
        assert np.array_equal(outs[""]["many"], outs[mode]["many"]), mode
    many = outs[""]["many"]
    assert np.isfinite(many).all() and np.abs(outs[""]["one"]).max() > 1e-3
    assert np.array_equal(many[0, : 131 * 256], outs[""]["one"])  # an item of a ragged batch == the same item alone
    assert not many[1, 47 * 256:].any()  # nothing behind an item's length


def test_encoder_bf16_mode_is_opt_in_bounded_and_resets():
    """BASELINE config 3's reduced-precision encoder (b200tts_call_set_encoder_bf16): off unless asked for, a small bounded
    drift of logw / mel when on (bf16 activations in the encoder GEMMs, fp32 duration head), and switching it off again
    restores the default arithmetic bit for bit.  tools/half_report.py measures the alignment mismatch rate."""
    g = np.load(os.path.join(GOLDEN, "bench128.npz"))
    glow = common.product_glow()
    call = engine.CALLS.get(0)
    ids = g["ids"]
    eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
    st = {"noise_scale": float(g["noise_scale"]), "length_scale": float(g["length_scale"]), "eps": eps}

    def run(**extra):
        mel = glow.phonemes_to_mels(ids, dict(st, **extra))
        logw, w_ceil, _ = glow.model.alignment(call, 1, len(ids), mel.shape[2])
        return mel, logw[0].copy(), w_ceil[0].copy()

    mel0, logw0, w0 = run()
    assert np.array_equal(w0, g["w_ceil"])
    mel1, logw1, w1 = run(encoder_bf16=True)
    d = float(np.abs(logw1 - logw0).max())
    print("encoder_bf16: max|dlogw| %.3e, w_ceil mismatches %d / %d" % (d, int((w1 != w0).sum()), len(ids)))
    assert 0.0 < d < 0.1  # the mode really changes the arithmetic, and only by bf16 rounding
    assert int((w1 != w0).sum()) <= len(ids) // 4
    if np.array_equal(w1, w0):
        assert float(np.abs(mel1 - mel0).max()) < 0.1
    mel2, logw2, w2 = run()  # the default again
    assert np.array_equal(logw2, logw0) and np.array_equal(mel2, mel0)
    glow.encoder_bf16 = True  # instance default, overridable per call
    try:
        _, logw3, _ = run()
        assert np.array_equal(logw3, logw1)
        _, logw4, _ = run(encoder_bf16=False)
        assert np.array_equal(logw4, logw0)
    finally:
        glow.encoder_bf16 = False


def test_device_resident_buffers():
    """`value` in bench.py is measured with inputs/outputs resident in HBM: same results as host buffers."""
    glow, voc = common.product_glow(), common.product_hifi()
    call = engine.CALLS.get(0)
    ids = synth.bench_ids(48, seed=8)
    eps = synth.seeded_eps(512, seed=2)
    mel_h = glow.phonemes_to_mels(ids, {"eps": eps})
    ty = mel_h.shape[2]
    dev = torch.device("cuda:0")
    ids_d = torch.from_numpy(ids).view(1, -1).to(dev)
    xlen_d = torch.tensor([len(ids)], dtype=torch.int32, device=dev)
    y = glow.model.encode(call, ids_d, xlen_d, 1, len(ids), 1.0)
    assert int(y[0]) == ty
    eps_d = torch.from_numpy(np.ascontiguousarray(eps[:, :ty])).to(dev)
    mel_d = torch.empty(1, 80, ty, device=dev)
    glow.model.decode(call, mel_d, ty, 0.667, eps_d, ty, 0)
    torch.cuda.synchronize()
    assert np.array_equal(mel_d.cpu().numpy(), mel_h)


def test_batch_scheduler_matches_per_sentence_results():
    """Length-bucketed micro-batches on two host threads return, in submission order, what each sentence gives alone."""
    from larynx_b200.scheduler import BatchScheduler

    glow, voc = common.product_glow(), common.product_hifi()
    glue = host.MelGlue.from_config(common.glow_cfg()["audio"])
    rs = np.random.RandomState(9)
    rows = [synth.bench_ids(int(n), seed=500 + i) for i, n in enumerate(rs.randint(0, 60, 13))]
    sch = BatchScheduler(glow, voc, glue, max_batch=4, max_cost=200, workers=2, noise_scale=0.0)
    out = sch.synthesize(rows)
    streamed = dict(sch.stream(rows))
    assert len(out) == len(rows) and sorted(streamed) == list(range(len(rows)))
    for i, r in enumerate(rows):
        if len(r) == 0:
            assert out[i].shape == (0,)
            continue
        one = host.synthesize_batch(glow, voc, glue, [r], noise_scale=0.0, want_float=True)
        assert out[i].dtype == np.int16 and out[i].shape == one["audio"][0].shape, i
        tol = _i16_tol(one["pcm"][0])
        assert np.abs(out[i].astype(np.int32) - one["audio"][0].astype(np.int32)).max() <= tol, i
        assert np.array_equal(streamed[i], out[i]), i  # same plan, same seeds, noise off: identical runs


@pytest.mark.parametrize("over", [
    dict(signal_norm=True, symmetric_norm=True, clip_norm=True, spec_gain=20.0),
    dict(signal_norm=True, symmetric_norm=False, clip_norm=True, spec_gain=20.0),
    dict(signal_norm=False, convert_db_to_amp=True, do_dynamic_range_compression=False, spec_gain=1.0),
    dict(signal_norm=False, convert_db_to_amp=False, do_dynamic_range_compression=True),
])
def test_device_glue_variants_match_host_glue(over):
    """The glue inside b200tts_tts_finish (struct b200tts_glue) against the numpy glue of the two-call path
    (larynx/__init__.py:242-249) for AudioSettings other than ljspeech's."""
    glow, voc = common.product_glow(), common.product_hifi()
    cfg = dict(common.glow_cfg()["audio"])
    cfg.update(max_norm=4.0, min_level_db=-100.0, ref_level_db=20.0)
    cfg.update(over)
    glue = host.MelGlue.from_config(cfg)
    ids = synth.bench_ids(24, seed=77)
    eps = synth.seeded_eps(512, seed=3)
    f = host.synthesize_batch(glow, voc, glue, [ids], 0.667, 1.0, eps=eps[None], want_float=True, want_mel=True)
    two_call = voc.mels_to_float(glue.apply(f["mel"][0][None]))
    assert f["pcm"][0].shape == two_call.shape
    assert np.all(np.isfinite(two_call))
    assert np.abs(f["pcm"][0] - two_call).max() <= TOL_E2E
