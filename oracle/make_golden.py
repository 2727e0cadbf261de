"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden

What is executed: the reference's own loaders and modules — glow_tts/checkpoint.py:26-68,
hifi_gan/checkpoint.py:36-70, FlowGenerator.forward (glow_tts/models.py:308-354) with
decoder.store_inverse() (larynx/glow_tts.py:94), Generator.forward (hifi_gan/models.py:186-202)
after remove_weight_norm() (larynx/hifi_gan.py:100), AudioSettings glue (larynx/audio.py:58-108)
and audio_float_to_int16 (larynx/audio.py:118-125) — on the synthetic reference-format voices of
larynx_b200/synth.py (seeds below).  Weights are regenerated bit-identically from the seeds on any
machine, so only inputs and outputs are committed.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from larynx_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

GLOW_SEED, HIFI_SEED, EPS_SEED = 1234, 4321, 1234
NOISE_SCALE, LENGTH_SCALE = 0.667, 1.0

FIXTURE_28 = "3 8 4 14 3 35 3 26 4 34 22 3 1 3 19 4 32 23 3 35 19 3 4 37 16 20 3 2"  # be_a_voice_not_an_echo


def run_case(glow, gen, ids, audio_settings, taps_at=()):
    import larynx.audio as laudio  # the reference's own glue

    eps = synth.seeded_eps(4096, seed=EPS_SEED)
    captured = {}
    hooks = []
    for blk in taps_at:
        # flows = [ActNorm_0, InvConv_0, Coupling_0, ...]; in reverse the ActNorm of block `blk` runs last
        mod = glow.decoder.flows[3 * blk]
        hooks.append(mod.register_forward_hook(lambda m, i, o, blk=blk: captured.__setitem__(blk, o[0].detach().clone())))
    r = ref_harness.run_reference_glow(glow, ids, NOISE_SCALE, LENGTH_SCALE, eps)
    for h in hooks:
        h.remove()
    mel = r["mel"].numpy()
    attn = r["attn"].squeeze(1)[0].numpy()  # [t_x, t_y] one-hot
    idx = attn.argmax(0).astype(np.int32)
    assert np.all(attn.sum(0) == 1.0)
    logw = r["logw"].numpy().reshape(-1)
    w = np.exp(logw)
    aset = laudio.AudioSettings(**audio_settings)
    m2 = mel
    if aset.signal_norm:
        m2 = aset.denormalize(m2)
    if aset.convert_db_to_amp:
        m2 = aset.db_to_amp(m2)
    if aset.do_dynamic_range_compression:
        m2 = aset.dynamic_range_compression(m2)
    m2 = m2.astype(np.float32)
    with torch.no_grad():
        pcm = gen(torch.from_numpy(m2)).squeeze(0).numpy()
    audio = laudio.audio_float_to_int16(pcm).squeeze()
    out = {
        "ids": np.asarray(ids, dtype=np.int64),
        "mel": mel[0],
        "mel_glued": m2[0],
        "logw": logw.astype(np.float32),
        "w_ceil": np.ceil(w * LENGTH_SCALE).astype(np.int32),
        "idx": idx,
        "pcm": pcm.reshape(-1).astype(np.float32),
        "audio": audio.astype(np.int16),
        "ceil_margin": np.array(np.abs(w - np.round(w)).min(), dtype=np.float32),
    }
    for blk, t in captured.items():
        out[f"flow{blk}"] = t[0].numpy()
    return out


def main():
    assert ref_harness.available(), "/root/reference is required to regenerate the golden vectors"
    tmp = Path(tempfile.mkdtemp())
    synth.write_glow_voice(tmp / "glow", 46, GLOW_SEED)
    synth.write_hifigan_voice(tmp / "hifi", "universal_large", HIFI_SEED)
    glow, gen = ref_harness.load_reference_models(tmp / "glow", tmp / "hifi")
    audio_settings = synth.glow_config(46)["audio"]
    outdir = ROOT / "tests" / "golden"
    outdir.mkdir(parents=True, exist_ok=True)
    meta = dict(glow_seed=GLOW_SEED, hifi_seed=HIFI_SEED, eps_seed=EPS_SEED, noise_scale=NOISE_SCALE, length_scale=LENGTH_SCALE)

    a = run_case(glow, gen, np.array([int(x) for x in FIXTURE_28.split()]), audio_settings, taps_at=(11, 6, 0))
    np.savez_compressed(outdir / "ljspeech_fixture28.npz", **a, **{k: np.array(v) for k, v in meta.items()})
    print("fixture28: T_y", a["mel"].shape[1], "ceil margin", float(a["ceil_margin"]))

    b = run_case(glow, gen, synth.bench_ids(128), audio_settings)
    b["pcm"] = b["pcm"][::16].copy()  # keep the fixture small: every 16th float sample + all int16 samples
    np.savez_compressed(outdir / "bench128.npz", **b, **{k: np.array(v) for k, v in meta.items()})
    print("bench128: T_y", b["mel"].shape[1], "ceil margin", float(b["ceil_margin"]))

    # second vocoder family (ResBlock2, 3 upsamplers): vctk_small on the fixture-28 mel
    synth.write_hifigan_voice(tmp / "hifi_small", "vctk_small", HIFI_SEED)
    synth.write_hifigan_voice(tmp / "hifi_medium", "vctk_medium", HIFI_SEED)
    for name in ("vctk_small", "vctk_medium"):
        _, g2 = ref_harness.load_reference_models(tmp / "glow", tmp / ("hifi_small" if name == "vctk_small" else "hifi_medium"))
        with torch.no_grad():
            pcm = g2(torch.from_numpy(a["mel_glued"][None])).reshape(-1).numpy()
        np.savez_compressed(outdir / f"{name}_fixture28.npz", mel_glued=a["mel_glued"], pcm=pcm.astype(np.float32),
                            hifi_seed=np.array(HIFI_SEED))
        print(name, "pcm std", float(pcm.std()), "absmax", float(np.abs(pcm).max()))


    make_denoiser(gen, a["mel_glued"], outdir)


def make_round2(outdir):
    """Round-2 additions: a voice whose symbol table is not ljspeech's (de-de: num_symbols 54, local/de-de/*-glow_tts/
    config.json) through the whole reference chain, and the two smaller vocoders on the 128-id benchmark mel."""
    tmp = Path(tempfile.mkdtemp())
    meta = dict(hifi_seed=HIFI_SEED, eps_seed=EPS_SEED, noise_scale=NOISE_SCALE, length_scale=LENGTH_SCALE)
    synth.write_glow_voice(tmp / "glow54", 54, GLOW_SEED + 1)
    synth.write_hifigan_voice(tmp / "hifi", "universal_large", HIFI_SEED)
    glow, gen = ref_harness.load_reference_models(tmp / "glow54", tmp / "hifi")
    ids = synth.bench_ids(47, 54, seed=54)
    assert ids.max() >= 46, "the vector must use symbols beyond ljspeech's table"
    c = run_case(glow, gen, ids, synth.glow_config(54)["audio"], taps_at=(0,))
    c["pcm"] = c["pcm"][::8].copy()
    np.savez_compressed(outdir / "de54_ids47.npz", **c, glow_seed=np.array(GLOW_SEED + 1), num_symbols=np.array(54),
                        **{k: np.array(v) for k, v in meta.items()})
    print("de54_ids47: T_y", c["mel"].shape[1], "ceil margin", float(c["ceil_margin"]), "max id", int(ids.max()))
    b = np.load(outdir / "bench128.npz")
    for name in ("vctk_small", "vctk_medium"):
        synth.write_hifigan_voice(tmp / name, name, HIFI_SEED)
        _, g2 = ref_harness.load_reference_models(tmp / "glow54", tmp / name)
        with torch.no_grad():
            pcm = g2(torch.from_numpy(b["mel_glued"][None])).reshape(-1).numpy()
        np.savez_compressed(outdir / f"{name}_bench128.npz", pcm=pcm[::16].astype(np.float32).copy(), hifi_seed=np.array(HIFI_SEED),
                            n_samples=np.array(pcm.shape[0]))
        print(name, "bench128 pcm std", float(pcm.std()), "absmax", float(np.abs(pcm).max()), "samples", pcm.shape[0])


def make_denoiser(gen, mel_glued, outdir):
    """Golden vectors of the spectral-subtraction denoiser made by the reference's own numpy STFT code
    (larynx/audio.py:232-306) in the order larynx/hifi_gan.py:158-203 applies it."""
    ref_harness.import_reference()
    import larynx.audio as la

    # (1) pure STFT arithmetic on random input with a strong bias (many clipped bins)
    rs = np.random.RandomState(3)
    audio = (rs.randn(1, 256 * 41) * 0.3).astype(np.float32)
    bias = np.abs(rs.randn(1, 513, 1)) * 0.5
    m, p = la.transform(audio)
    out = la.inverse(np.clip(m - bias * 0.7, a_min=0.0, a_max=None), p)
    # (2) the CLI default on the fixture-28 utterance: bias from the vocoder on zeros[1, 80, 88], strength 0.005
    with torch.no_grad():
        bias_audio = gen(torch.zeros(1, 80, 88)).squeeze(0).numpy()
        pcm = gen(torch.from_numpy(mel_glued[None])).squeeze(0).numpy()
    bias_spec, _ = la.transform(bias_audio)
    bias_spec = bias_spec[:, :, 0][:, :, None]
    m, p = la.transform(pcm)
    pcm_dn = la.inverse(np.clip(m - bias_spec * 0.005, a_min=0.0, a_max=None), p)
    np.savez_compressed(outdir / "denoiser.npz", audio=audio, bias=bias, strength=np.array(0.7), out=out,
                        fixture28_bias_spec=bias_spec.astype(np.float32), fixture28_strength=np.array(0.005),
                        fixture28_pcm_denoised=pcm_dn.astype(np.float32).reshape(-1),
                        fixture28_audio_denoised=la.audio_float_to_int16(pcm_dn).squeeze())
    print("denoiser: bias max", float(bias_spec.max()), "pcm_dn absmax", float(np.abs(pcm_dn).max()))


if __name__ == "__main__":
    import sys

    if "--round2" in sys.argv:
        make_round2(ROOT / "tests" / "golden")
    elif "--denoiser-only" in sys.argv:
        _tmp = Path(tempfile.mkdtemp())
        synth.write_glow_voice(_tmp / "glow", 46, GLOW_SEED)
        synth.write_hifigan_voice(_tmp / "hifi", "universal_large", HIFI_SEED)
        _, _gen = ref_harness.load_reference_models(_tmp / "glow", _tmp / "hifi")
        _g = np.load(ROOT / "tests" / "golden" / "ljspeech_fixture28.npz")
        make_denoiser(_gen, _g["mel_glued"], ROOT / "tests" / "golden")
    else:
        main()
