"""Synthetic voices in the reference's on-disk format.

The reference tree ships no model weights (they are downloaded at run time,
larynx/utils.py:19-21, 104-167), and there is no network here, so parity tests
and the benchmark use deterministic *synthetic* weights of the exact reference
architecture.  This module writes them as state-dicts with the reference's own
key names and shapes, i.e. what `glow_tts/checkpoint.py:26-68` reads from
`generator.pth["model"]` and `hifi_gan/checkpoint.py:36-70` reads from
`generator.pth["generator"]` (weight-norm still un-folded: `weight_g/weight_v`).

Every tensor is re-randomised — the reference's default init zeroes
`CouplingBlock.end` (glow_tts/attentions.py:104-106) and the prenet `proj`
(glow_tts/layers.py:69-71), and HiFi-GAN's N(0, 0.01) init gives |audio|~3e-3,
which would make a 1e-3 tolerance vacuous.  Scales are
chosen so every intermediate stays O(1).

Random numbers come from numpy's legacy `RandomState` (stream is stable across
numpy versions and machines), so the build container and the GPU box generate
bit-identical weights from a seed and the committed golden vectors stay valid.
"""
from __future__ import annotations

import json
import math
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from .constants import HIFIGAN_CONFIGS  # noqa: F401  (re-exported for tests / bench)

# Architecture constants shared by all 51 shipped voices
# (local/en-us/ljspeech-glow_tts/config.json:36-62).
GLOW_DEFAULT_MODEL = {
    "num_symbols": 46,
    "hidden_channels": 192,
    "filter_channels": 768,
    "filter_channels_dp": 256,
    "kernel_size": 3,
    "p_dropout": 0.1,
    "n_blocks_dec": 12,
    "n_layers_enc": 6,
    "n_heads": 2,
    "p_dropout_dec": 0.05,
    "dilation_rate": 1,
    "kernel_size_dec": 5,
    "n_block_layers": 4,
    "n_sqz": 2,
    "prenet": True,
    "mean_only": True,
    "hidden_channels_enc": 192,
    "hidden_channels_dec": 192,
    "window_size": 4,
    "n_speakers": 1,
    "n_split": 4,
    "sigmoid_scale": False,
    "block_length": None,
    "gin_channels": 0,
    "n_frames_per_step": 1,
}

GLOW_DEFAULT_AUDIO = {
    "filter_length": 1024,
    "hop_length": 256,
    "win_length": 1024,
    "mel_channels": 80,
    "sample_rate": 22050,
    "sample_bytes": 2,
    "channels": 1,
    "mel_fmin": 0.0,
    "mel_fmax": 8000.0,
    "ref_level_db": 20.0,
    "spec_gain": 1.0,
    "signal_norm": True,
    "min_level_db": -100.0,
    "max_norm": 1.0,
    "clip_norm": True,
    "symmetric_norm": True,
    "do_dynamic_range_compression": True,
    "convert_db_to_amp": True,
}

# hifi_gan/config.py:29-40 defaults == local/hifi_gan/universal_large/config.json:2,10-14
def glow_config(num_symbols: int = 46, **model_overrides) -> dict:
    """A full reference-format GlowTTS config.json dict."""
    model = dict(GLOW_DEFAULT_MODEL)
    model["num_symbols"] = int(num_symbols)
    model.update(model_overrides)
    return {"seed": 1234, "audio": dict(GLOW_DEFAULT_AUDIO), "model": model, "version": 1}


def hifigan_config(name: str = "universal_large") -> dict:
    """Nested (dataclasses-json) form of a vocoder config: {"model": {...}, "audio": {...}}."""
    return {
        "seed": 1234,
        "audio": {"num_mels": 80, "sampling_rate": 22050, "hop_length": 256},
        "model": dict(HIFIGAN_CONFIGS[name]),
        "version": 1,
    }


class _Rng:
    def __init__(self, seed: int):
        self.rs = np.random.RandomState(int(seed))

    def normal(self, shape, std=1.0):
        return (self.rs.standard_normal(size=shape) * std).astype(np.float32)

    def uniform(self, shape, lo, hi):
        return self.rs.uniform(lo, hi, size=shape).astype(np.float32)


def _wn(sd: Dict[str, np.ndarray], prefix: str, v: np.ndarray, rng: _Rng, bias_std=0.02):
    """Store a weight-normed conv: weight_v = v, weight_g = ||v|| * U(0.8, 1.2) per dim-0 slice.

    torch.nn.utils.weight_norm uses dim=0 for Conv1d *and* ConvTranspose1d, so
    `weight_g` has shape (v.shape[0], 1, 1) in both cases.
    """
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
    g = norm * rng.uniform(norm.shape, 0.8, 1.2)
    sd[prefix + ".weight_g"] = g.astype(np.float32)
    sd[prefix + ".weight_v"] = v
    return g


def glow_state_dict(config: dict, seed: int = 1234) -> Dict[str, np.ndarray]:
    """State-dict (numpy) for glow_tts.models.FlowGenerator built from `config`."""
    m = config["model"]
    assert m["n_speakers"] <= 1 and m["gin_channels"] == 0, "multi-speaker voices are out of scope"
    assert m["mean_only"] and m["prenet"] and not m["sigmoid_scale"]
    rng = _Rng(seed)
    sd: Dict[str, np.ndarray] = {}
    H = m["hidden_channels_enc"] or m["hidden_channels"]
    Hd = m["hidden_channels_dec"] or m["hidden_channels"]
    F = m["filter_channels"]
    Fd = m["filter_channels_dp"]
    K = m["kernel_size"]
    mel = config["audio"]["mel_channels"]
    nh = m["n_heads"]
    dk = H // nh
    ws = m["window_size"]

    def ln(prefix, c):
        sd[prefix + ".gamma"] = (1.0 + rng.normal((c,), 0.1)).astype(np.float32)
        sd[prefix + ".beta"] = rng.normal((c,), 0.1)

    def conv(prefix, co, ci, k, gain=1.0, bias_std=0.02):
        sd[prefix + ".weight"] = rng.normal((co, ci, k), gain / math.sqrt(ci * k))
        sd[prefix + ".bias"] = rng.normal((co,), bias_std)

    # --- TextEncoder (glow_tts/models.py:52-140)
    sd["encoder.emb.weight"] = rng.normal((m["num_symbols"], H), H ** -0.5)
    for i in range(3):
        conv(f"encoder.pre.conv_layers.{i}", H, H, 5, gain=1.0)
    for i in range(3):
        ln(f"encoder.pre.norm_layers.{i}", H)
    conv("encoder.pre.proj", H, H, 1, gain=0.7)
    for i in range(m["n_layers_enc"]):
        p = f"encoder.encoder.attn_layers.{i}"
        sd[p + ".emb_rel_k"] = rng.normal((1, 2 * ws + 1, dk), dk ** -0.5)
        sd[p + ".emb_rel_v"] = rng.normal((1, 2 * ws + 1, dk), dk ** -0.5)
        conv(p + ".conv_q", H, H, 1, gain=1.2)
        conv(p + ".conv_k", H, H, 1, gain=1.2)
        conv(p + ".conv_v", H, H, 1)
        conv(p + ".conv_o", H, H, 1)
    for i in range(m["n_layers_enc"]):
        ln(f"encoder.encoder.norm_layers_1.{i}", H)
    for i in range(m["n_layers_enc"]):
        conv(f"encoder.encoder.ffn_layers.{i}.conv_1", F, H, K, gain=math.sqrt(2.0))
        conv(f"encoder.encoder.ffn_layers.{i}.conv_2", H, F, K, gain=1.0)
    for i in range(m["n_layers_enc"]):
        ln(f"encoder.encoder.norm_layers_2.{i}", H)
    conv("encoder.proj_m", mel, H, 1, gain=0.8)
    conv("encoder.proj_w.conv_1", Fd, H, K, gain=math.sqrt(2.0))
    ln("encoder.proj_w.norm_1", Fd)
    conv("encoder.proj_w.conv_2", Fd, Fd, K, gain=math.sqrt(2.0))
    ln("encoder.proj_w.norm_2", Fd)
    # duration head: logw = ln 4 + small  =>  w = exp(logw) in roughly [1.5, 9]
    sd["encoder.proj_w.proj.weight"] = rng.normal((1, Fd, 1), 1.5 / math.sqrt(Fd))
    sd["encoder.proj_w.proj.bias"] = np.array([0.06], dtype=np.float32)

    # --- FlowSpecDecoder (glow_tts/models.py:143-213): flows = [ActNorm, InvConvNear, CouplingBlock] * n_blocks
    C = mel * m["n_sqz"]
    ns = m["n_split"]
    kd = m["kernel_size_dec"]
    nl = m["n_block_layers"]
    for b in range(m["n_blocks_dec"]):
        a, inv, cp = 3 * b, 3 * b + 1, 3 * b + 2
        if b > 0:
            # inverse ActNorm multiplies by exp(-logs): a mild per-block contraction keeps the
            # 12-block reverse pass O(1) (the coupling's "- m" term adds variance every block)
            sd[f"decoder.flows.{a}.logs"] = (0.18 + rng.normal((1, C, 1), 0.1)).astype(np.float32)
            sd[f"decoder.flows.{a}.bias"] = rng.normal((1, C, 1), 0.1)
        else:
            # flows[0] is applied LAST in reverse: squash to the range real normalised mels
            # occupy (~0.56 +- 0.04, i.e. log10-amplitude of about -2 +- 2 after denormalize,
            # larynx/audio.py:83-104) so the host glue's clip/exp/log is exercised un-saturated
            logs0 = (2.5 + rng.normal((1, C, 1), 0.1)).astype(np.float32)
            sd[f"decoder.flows.{a}.logs"] = logs0
            sd[f"decoder.flows.{a}.bias"] = (-(0.56 + rng.normal((1, C, 1), 0.01)) * np.exp(logs0)).astype(np.float32)
        q, r = np.linalg.qr(rng.rs.standard_normal(size=(ns, ns)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        # not exactly orthogonal: perturb so that W^-1 != W^T
        q = q + rng.rs.standard_normal(size=(ns, ns)) * 0.05
        sd[f"decoder.flows.{inv}.weight"] = q.astype(np.float32)
        p = f"decoder.flows.{cp}"
        sd[p + ".start.bias"] = rng.normal((Hd,), 0.02)
        _wn(sd, p + ".start", rng.normal((Hd, C // 2, 1), 1.0 / math.sqrt(C // 2)), rng)
        sd[p + ".end.weight"] = rng.normal((C, Hd, 1), 0.25 / math.sqrt(Hd))
        sd[p + ".end.bias"] = rng.normal((C,), 0.02)
        for l in range(nl):
            sd[p + f".wn.in_layers.{l}.bias"] = rng.normal((2 * Hd,), 0.02)
            _wn(sd, p + f".wn.in_layers.{l}", rng.normal((2 * Hd, Hd, kd), 1.3 / math.sqrt(Hd * kd)), rng)
        for l in range(nl):
            rc = 2 * Hd if l < nl - 1 else Hd
            sd[p + f".wn.res_skip_layers.{l}.bias"] = rng.normal((rc,), 0.02)
            _wn(sd, p + f".wn.res_skip_layers.{l}", rng.normal((rc, Hd, 1), 1.6 / math.sqrt(Hd)), rng)
    return sd


def hifigan_state_dict(config: dict, seed: int = 4321) -> Dict[str, np.ndarray]:
    """State-dict (numpy) for hifi_gan.models.Generator (hifi_gan/models.py:148-211)."""
    m = config["model"]
    rng = _Rng(seed)
    sd: Dict[str, np.ndarray] = {}
    C0 = m["upsample_initial_channel"]
    lgain = math.sqrt(2.0 / 1.01)  # variance-preserving gain in front of LeakyReLU(0.1)

    sd["conv_pre.bias"] = rng.normal((C0,), 0.02)
    _wn(sd, "conv_pre", rng.normal((C0, 80, 7), 0.25 / math.sqrt(80 * 7)), rng)
    ch = C0
    for i, (u, k) in enumerate(zip(m["upsample_rates"], m["upsample_kernel_sizes"])):
        cin, cout = C0 // (2 ** i), C0 // (2 ** (i + 1))
        sd[f"ups.{i}.bias"] = rng.normal((cout,), 0.02)
        # every output sample sees k/u taps of cin channels
        _wn(sd, f"ups.{i}", rng.normal((cin, cout, k), lgain / math.sqrt(cin * k / u)), rng)
        ch = cout
    n = 0
    for i in range(len(m["upsample_rates"])):
        ch = C0 // (2 ** (i + 1))
        for k, dil in zip(m["resblock_kernel_sizes"], m["resblock_dilation_sizes"]):
            p = f"resblocks.{n}"
            if m["resblock"] == "1":
                for j in range(len(dil)):
                    sd[p + f".convs1.{j}.bias"] = rng.normal((ch,), 0.02)
                    _wn(sd, p + f".convs1.{j}", rng.normal((ch, ch, k), lgain / math.sqrt(ch * k)), rng)
                for j in range(len(dil)):
                    sd[p + f".convs2.{j}.bias"] = rng.normal((ch,), 0.02)
                    _wn(sd, p + f".convs2.{j}", rng.normal((ch, ch, k), 0.45 * lgain / math.sqrt(ch * k)), rng)
            else:
                for j in range(len(dil)):
                    sd[p + f".convs.{j}.bias"] = rng.normal((ch,), 0.02)
                    _wn(sd, p + f".convs.{j}", rng.normal((ch, ch, k), 0.6 * lgain / math.sqrt(ch * k)), rng)
            n += 1
    sd["conv_post.bias"] = rng.normal((1,), 0.02)
    _wn(sd, "conv_post", rng.normal((1, ch, 7), 0.55 / math.sqrt(ch * 7)), rng)
    return sd


# -----------------------------------------------------------------------------
# Reference-format voice directories (what larynx.load_tts_model / load_vocoder_model read)
# -----------------------------------------------------------------------------


def write_glow_voice(path: Path, num_symbols: int = 46, seed: int = 1234, config: Optional[dict] = None) -> Path:
    """Write `<path>/{config.json, generator.pth}`; generator.pth = {"model": sd, "version": 1}."""
    import torch

    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    cfg = config or glow_config(num_symbols)
    sd = glow_state_dict(cfg, seed)
    with open(path / "config.json", "w", encoding="utf-8") as f:
        json.dump(cfg, f, indent=4)
    torch.save({"model": {k: torch.from_numpy(v) for k, v in sd.items()}, "version": 1}, str(path / "generator.pth"))
    return path


def write_hifigan_voice(path: Path, name: str = "universal_large", seed: int = 4321) -> Path:
    """Write `<path>/{config.json, generator.pth}`; generator.pth = {"generator": sd}."""
    import torch

    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    cfg = hifigan_config(name)
    sd = hifigan_state_dict(cfg, seed)
    with open(path / "config.json", "w", encoding="utf-8") as f:
        json.dump(cfg, f, indent=4)
    torch.save({"generator": {k: torch.from_numpy(v) for k, v in sd.items()}}, str(path / "generator.pth"))
    return path


def bench_ids(n: int = 128, num_symbols: int = 46, seed: int = 1234) -> np.ndarray:
    """The canonical benchmark input: RandomState(1234).randint(1, 46, 128)."""
    return np.random.RandomState(seed).randint(1, num_symbols, n).astype(np.int64)


def seeded_eps(n_frames: int, mel_channels: int = 80, seed: int = 1234) -> np.ndarray:
    """Injected noise eps[80, T]: column t depends only on (seed, t) so any prefix is reproducible."""
    rs = np.random.RandomState(seed)
    return np.ascontiguousarray(rs.standard_normal(size=(n_frames, mel_channels)).astype(np.float32).T)
