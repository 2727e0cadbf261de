"""Pins oracle/port.py (the CPU restatement) against the reference.

(a) everywhere: against tests/golden/*.npz, produced by oracle/make_golden.py from the UNMODIFIED
    reference modules run in the build container;
(b) where /root/reference exists: live against the reference's own classes, including the skew-trick
    attention (glow_tts/attentions.py:284-335) that the port restates in closed form.
"""
import os

import numpy as np
import pytest
import torch

from larynx_b200 import synth
from oracle import port, ref_harness
from tests import common

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("name", ["ljspeech_fixture28.npz", "bench128.npz"])
def test_port_matches_golden(name):
    g = _load(name)
    eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
    o = port.sentence_task(common.oracle_glow_sd(), common.glow_cfg(), common.oracle_hifi_sd(), common.hifi_cfg(),
                           g["ids"], float(g["noise_scale"]), float(g["length_scale"]), eps)
    # alignment: bit exact
    assert np.array_equal(o["w_ceil"][0].astype(np.int32), g["w_ceil"])
    assert np.array_equal(o["idx"][0].astype(np.int32), g["idx"])
    assert o["mel"].shape[2] == g["mel"].shape[1]
    assert np.abs(o["logw"].reshape(-1) - g["logw"]).max() <= 1e-6
    assert np.abs(o["mel"][0] - g["mel"]).max() <= 5e-6
    assert np.abs(o["mel_glued"][0] - g["mel_glued"]).max() <= 1e-3  # glue multiplies mel error by 50*ln(10)
    # vocoder pinned on the golden glued mel (so the comparison is not diluted by mel differences)
    pcm = port.hifigan_forward(common.oracle_hifi_sd(), common.hifi_cfg(), torch.from_numpy(g["mel_glued"][None])).reshape(-1).numpy()
    ref = g["pcm"]
    got = pcm if len(ref) == len(pcm) else pcm[::16]
    assert np.abs(got - ref).max() <= 2e-5
    assert np.abs(port.float_to_int16(pcm).astype(np.int32) - g["audio"].astype(np.int32)).max() <= 1


@pytest.mark.parametrize("name", ["vctk_small", "vctk_medium"])
def test_port_vocoder_variants_match_golden(name):
    g = _load(f"{name}_fixture28.npz")
    pcm = port.hifigan_forward(common.oracle_hifi_sd(name), common.hifi_cfg(name), torch.from_numpy(g["mel_glued"][None]))
    assert np.abs(pcm.reshape(-1).numpy() - g["pcm"]).max() <= 2e-5


def test_port_matches_golden_other_symbol_table():
    """A voice that is not ljspeech's shape: num_symbols 54 (de-de), own weights, ids beyond ljspeech's table."""
    g = _load("de54_ids47.npz")
    ns = int(g["num_symbols"])
    assert ns == 54 and g["ids"].max() >= 46
    eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
    gsd, gcfg = common.oracle_glow_sd(ns, int(g["glow_seed"])), common.glow_cfg(ns)
    o = port.sentence_task(gsd, gcfg, common.oracle_hifi_sd(), common.hifi_cfg(), g["ids"], float(g["noise_scale"]),
                           float(g["length_scale"]), eps)
    assert float(g["ceil_margin"]) > 1e-4
    assert np.array_equal(o["w_ceil"][0].astype(np.int32), g["w_ceil"])
    assert np.array_equal(o["idx"][0].astype(np.int32), g["idx"])
    assert np.abs(o["logw"].reshape(-1) - g["logw"]).max() <= 1e-6
    assert np.abs(o["mel"][0] - g["mel"]).max() <= 5e-6
    pcm = port.hifigan_forward(common.oracle_hifi_sd(), common.hifi_cfg(), torch.from_numpy(g["mel_glued"][None])).reshape(-1).numpy()
    assert np.abs(pcm[::8] - g["pcm"]).max() <= 2e-5
    assert np.abs(port.float_to_int16(pcm).astype(np.int32) - g["audio"].astype(np.int32)).max() <= 1


@pytest.mark.parametrize("name", ["vctk_small", "vctk_medium"])
def test_port_vocoder_variants_match_golden_at_bench_size(name):
    g, b = _load(f"{name}_bench128.npz"), _load("bench128.npz")
    pcm = port.hifigan_forward(common.oracle_hifi_sd(name), common.hifi_cfg(name), torch.from_numpy(b["mel_glued"][None]))
    pcm = pcm.reshape(-1).numpy()
    assert pcm.shape[0] == int(g["n_samples"])
    assert np.abs(pcm[::16] - g["pcm"]).max() <= 2e-5


def test_flow_taps_match_golden():
    g = _load("ljspeech_fixture28.npz")
    eps = synth.seeded_eps(4096, seed=int(g["eps_seed"]))
    taps = {}
    ids = torch.from_numpy(g["ids"]).unsqueeze(0)
    port.glow_forward(common.oracle_glow_sd(), common.glow_cfg(), ids, torch.tensor([ids.shape[1]]), float(g["noise_scale"]),
                      float(g["length_scale"]), torch.from_numpy(eps).unsqueeze(0), taps)
    for blk in (11, 6, 0):
        assert np.abs(taps[f"flow{blk}"][0].numpy() - g[f"flow{blk}"]).max() <= 2e-5


def test_golden_vectors_have_ceil_margin():
    # "alignment bit-exact" is only meaningful when no duration sits on an integer boundary
    for name in ("ljspeech_fixture28.npz", "bench128.npz"):
        assert float(_load(name)["ceil_margin"]) > 1e-4


def test_glue_closed_form():
    # larynx/audio.py:83-108 with the ljspeech settings: ln(max(10^(50*clip(m,-1,1)-30), 1e-5))
    m = np.linspace(-1.5, 1.5, 301, dtype=np.float32).reshape(1, 1, -1)
    out = port.mel_glue(m, common.glow_cfg()["audio"])
    ref = np.log(10.0) * np.maximum(50.0 * np.clip(m.astype(np.float64), -1, 1) - 30.0, -5.0)
    assert np.abs(out - ref).max() < 2e-5 * np.abs(ref).max()


def test_int16_conversion_matches_reference_semantics():
    a = np.array([0.5, -0.25, 0.1234567, -0.9999], dtype=np.float32)
    out = port.float_to_int16(a)
    scale = 32767.0 / 0.9999
    assert out.dtype == np.int16
    assert np.array_equal(out, np.trunc(np.clip(a * np.float32(scale), -32767, 32767)).astype(np.int16))
    quiet = np.array([0.001, -0.002], dtype=np.float32)  # below the 0.01 floor: gain is capped
    assert np.array_equal(port.float_to_int16(quiet), np.trunc(quiet * np.float32(32767.0 / 0.01)).astype(np.int16))


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present on this machine")
def test_port_matches_live_reference(tmp_path):
    synth.write_glow_voice(tmp_path / "glow", 46, common.GLOW_SEED)
    synth.write_hifigan_voice(tmp_path / "hifi", "universal_large", common.HIFI_SEED)
    glow, gen = ref_harness.load_reference_models(tmp_path / "glow", tmp_path / "hifi")
    ids = common.fixture_ids("be_a_voice_not_an_echo")[:19]
    eps = synth.seeded_eps(512, seed=7)
    for ls in (1.0, 1.3):
        r = ref_harness.run_reference_glow(glow, ids, 0.5, ls, eps)
        o = port.sentence_task(common.oracle_glow_sd(), common.glow_cfg(), common.oracle_hifi_sd(), common.hifi_cfg(), ids, 0.5, ls, eps)
        assert o["mel"].shape == tuple(r["mel"].shape)
        assert np.abs(o["mel"] - r["mel"].numpy()).max() <= 5e-6
        assert np.array_equal(r["attn"].squeeze(1)[0].numpy().argmax(0), o["idx"][0])
        with torch.no_grad():
            pcm = gen(torch.from_numpy(o["mel_glued"])).reshape(-1).numpy()
        assert np.abs(pcm - o["pcm"]).max() <= 2e-5


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present on this machine")
def test_reference_batched_equals_single():
    # the padded-batch GlowTTS reproduces per-utterance mels on the valid region
    ids_a, ids_b = synth.bench_ids(20, seed=1), synth.bench_ids(33, seed=2)
    sd, cfg = common.oracle_glow_sd(), common.glow_cfg()
    x = torch.zeros(2, 33, dtype=torch.int64)
    x[0, :20] = torch.from_numpy(ids_a)
    x[1] = torch.from_numpy(ids_b)
    both = port.glow_forward(sd, cfg, x, torch.tensor([20, 33]), 0.0, 1.0)
    one = port.glow_forward(sd, cfg, torch.from_numpy(ids_a).unsqueeze(0), torch.tensor([20]), 0.0, 1.0)
    ty = int(one["y_lengths"][0])
    assert int(both["y_lengths"][0]) == ty
    assert np.abs(both["mel"][0, :, :ty].numpy() - one["mel"][0].numpy()).max() < 2e-5
    assert float(both["mel"][0, :, ty:].abs().max()) == 0.0


def test_denoiser_restatements_equal_reference():
    """Loop-form (oracle/port.py) and vectorised (oracle/denoise_np.py) restatements of the denoiser agree with each
    other, with the committed golden vector (made by the reference's own larynx.audio.transform / inverse) and with
    those functions themselves where the reference can be imported."""
    from oracle import denoise_np as denoise

    rs = np.random.RandomState(3)
    audio = (rs.randn(1, 256 * 41) * 0.3).astype(np.float32)
    bias = np.abs(rs.randn(1, 513, 1)) * 0.5
    a = denoise.denoise(audio, bias, 0.7)
    b = port.denoise_reference(audio, bias, 0.7)
    assert a.shape == b.shape == (1, 256 * 41)
    assert np.abs(a - b).max() < 1e-12
    assert np.array_equal(denoise.audio_float_to_int16(a), port.float_to_int16(b))
    assert denoise.denoise(np.zeros((1, 300), np.float32), bias, 0.5).shape == (1, 1024)  # shorter than one frame
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "denoiser.npz"))
    for fn in (denoise.denoise, port.denoise_reference):
        assert np.abs(fn(g["audio"], g["bias"], float(g["strength"])) - g["out"]).max() < 1e-12
    if ref_harness.available():
        ref_harness.import_reference()
        import larynx.audio as la

        m, p = la.transform(audio)
        r = la.inverse(np.clip(m - bias * 0.7, 0, None), p)
        assert np.abs(r - a).max() < 1e-12


def test_glue_variants_host_mirror_oracle_and_reference_agree():
    """Every AudioSettings combination the glue can take (larynx/audio.py:58-59,83-108): the product's host mirror
    (larynx_b200.host.MelGlue, used between the two boundary calls), the oracle restatement, and the reference's own
    AudioSettings where it can be imported."""
    from larynx_b200 import host

    m = np.random.RandomState(5).uniform(-5, 5, size=(1, 80, 37)).astype(np.float32)
    la = None
    if ref_harness.available():
        ref_harness.import_reference()
        import larynx.audio as la  # noqa: F811
    for signal_norm in (False, True):
        for symmetric in (True, False):
            for clip in (True, False):
                for db2amp, drc in ((True, True), (True, False), (False, True)):
                    cfg = dict(common.glow_cfg()["audio"])
                    cfg.update(signal_norm=signal_norm, symmetric_norm=symmetric, clip_norm=clip, convert_db_to_amp=db2amp,
                               do_dynamic_range_compression=drc, max_norm=4.0, min_level_db=-100.0, ref_level_db=20.0,
                               spec_gain=20.0 if signal_norm else 1.0)
                    x = m if signal_norm else m * 0.2  # keep 10**x finite without normalisation
                    a = host.MelGlue.from_config(cfg).apply(x)
                    b = port.mel_glue(x, cfg)
                    assert a.dtype == np.float32 and np.array_equal(a, b), cfg
                    if la is not None:
                        s = la.AudioSettings(**{k: v for k, v in cfg.items() if k in la.AudioSettings.__dataclass_fields__})
                        r = x
                        if s.signal_norm and not s.clip_norm:
                            # reference bug: denormalize() only binds its result inside `if self.clip_norm`
                            # (larynx/audio.py:87-91) and raises; no shipped voice has clip_norm false.  The product
                            # and the oracle treat it as "no clipping".
                            with pytest.raises(UnboundLocalError):
                                s.denormalize(r)
                            continue
                        if s.signal_norm:
                            r = s.denormalize(r)
                        if s.convert_db_to_amp:
                            r = s.db_to_amp(r)
                        if s.do_dynamic_range_compression:
                            r = s.dynamic_range_compression(r)
                        assert np.allclose(a, r, rtol=1e-6, atol=1e-6), cfg
