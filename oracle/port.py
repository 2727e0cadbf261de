"""TEST INFRASTRUCTURE — CPU restatement (oracle) of the Larynx hot path.

    phoneme ids -> GlowTTS forward -> mel -> (host mel glue) -> HiFi-GAN forward -> PCM

This file restates, in functional fp32 PyTorch-on-CPU form, what the reference's
in-tree modules compute (the ONNX graphs the reference runs by default were
exported from exactly those modules, README.md:427,490).  Every function cites the
reference file:line it follows.  It is *pinned*: tests/test_oracle_pinned.py checks
it (a) live against the reference's own classes when /root/reference is present and
(b) everywhere against tests/golden/*.npz, which oracle/make_golden.py produced by
running the unmodified reference modules in the build container.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (larynx_b200/) never does.

Operates directly on reference-format state-dicts (weight_g/weight_v un-folded),
so it also checks the product's weight-norm folding / InvConv pre-inversion.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def to_torch_sd(sd_np: Dict[str, np.ndarray]) -> SD:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}


# -----------------------------------------------------------------------------
# shared helpers
# -----------------------------------------------------------------------------


def fold_weight_norm(sd: SD, prefix: str) -> Tensor:
    """w = g * v / ||v||, norm over all dims but 0 — what torch's weight_norm hook computes
    (used un-removed by CouplingBlock.start, glow_tts/attentions.py:100-102, and removed
    at load by WN.remove_weight_norm glow_tts/layers.py:164-170 /
    Generator.remove_weight_norm hifi_gan/models.py:204-211)."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    v = sd[prefix + ".weight_v"]
    g = sd[prefix + ".weight_g"]
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / norm)


def prefold(sd: SD, keep_runtime=(".start",)) -> SD:
    """Fold weight-norm ONCE, as the reference does at load (WN.remove_weight_norm glow_tts/layers.py:164-170 via
    store_inverse, Generator.remove_weight_norm hifi_gan/models.py:204-211).  CouplingBlock.start keeps its weight-norm
    at run time in the reference (glow_tts/attentions.py:100-102), so it stays unfolded here too.  Timing runs use this."""
    out = dict(sd)
    for k in sd:
        if k.endswith(".weight_v"):
            prefix = k[: -len(".weight_v")]
            if not any(prefix.endswith(sfx) for sfx in keep_runtime):
                out[prefix + ".weight"] = fold_weight_norm(sd, prefix)
    return out


def channel_layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-4) -> Tensor:
    """glow_tts/layers.py:19-28 — LN over dim 1 (channels), biased variance, eps 1e-4."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def length_mask(lengths: Tensor, t_max: int) -> Tensor:
    """glow_tts/utils.py:52-56 — [b, t] bool, t < length."""
    return torch.arange(t_max, dtype=lengths.dtype).unsqueeze(0) < lengths.unsqueeze(1)


# -----------------------------------------------------------------------------
# GlowTTS
# -----------------------------------------------------------------------------


def rel_attention(q: Tensor, k: Tensor, v: Tensor, emb_k: Tensor, emb_v: Tensor, mask2d: Tensor,
                  n_heads: int, window: int) -> Tensor:
    """glow_tts/attentions.py:214-264 in closed form (no pad/reshape skew trick).

    scores[i,j] = (q_i.k_j + [|j-i|<=w] q_i.Ek[j-i+w]) / sqrt(dk); masked_fill(mask==0,-1e4);
    softmax over j; out_i = sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w].
    The skew helpers (:284-335) build exactly this banded gather (verified by the pin test).
    """
    b, d, t = q.shape
    dk = d // n_heads
    qh = q.view(b, n_heads, dk, t).transpose(2, 3)
    kh = k.view(b, n_heads, dk, t).transpose(2, 3)
    vh = v.view(b, n_heads, dk, t).transpose(2, 3)
    scores = torch.matmul(qh, kh.transpose(-2, -1)) / math.sqrt(dk)
    idx = torch.arange(t)
    rel = idx.unsqueeze(0) - idx.unsqueeze(1)  # j - i
    band = rel.abs() <= window
    rel_c = (rel + window).clamp(0, 2 * window)
    qe = torch.matmul(qh, emb_k[0].t())  # [b,h,t,2w+1]
    rel_logits = torch.gather(qe, 3, rel_c.expand(b, n_heads, t, t)) * band
    scores = scores + rel_logits / math.sqrt(dk)
    scores = scores.masked_fill(mask2d.unsqueeze(1) == 0, -1e4)
    p = F.softmax(scores, dim=-1)
    out = torch.matmul(p, vh)
    pb = p * band
    relw = torch.zeros(b, n_heads, t, 2 * window + 1)
    relw.scatter_add_(3, rel_c.expand(b, n_heads, t, t), pb)
    out = out + torch.matmul(relw, emb_v[0])
    return out.transpose(2, 3).contiguous().view(b, d, t)


def text_encoder(sd: SD, cfg: dict, ids: Tensor, x_lengths: Tensor, taps: Optional[dict] = None):
    """glow_tts/models.py:118-140 (TextEncoder.forward) with prenet layers.py:73-80,
    Encoder attentions.py:62-74, FFN attentions.py:375-383, DurationPredictor models.py:39-49."""
    m = cfg["model"]
    H = m["hidden_channels_enc"] or m["hidden_channels"]
    x = F.embedding(ids, sd["encoder.emb.weight"]) * math.sqrt(H)
    x = x.transpose(1, 2)
    x_mask = length_mask(x_lengths, x.size(2)).unsqueeze(1).to(x.dtype)
    # prenet: conv -> LN -> ReLU (x3), then 1x1 proj + residual
    x_org = x
    for i in range(3):
        x = F.conv1d(x * x_mask, sd[f"encoder.pre.conv_layers.{i}.weight"], sd[f"encoder.pre.conv_layers.{i}.bias"], padding=2)
        x = channel_layer_norm(x, sd[f"encoder.pre.norm_layers.{i}.gamma"], sd[f"encoder.pre.norm_layers.{i}.beta"])
        x = torch.relu(x)
    x = x_org + F.conv1d(x, sd["encoder.pre.proj.weight"], sd["encoder.pre.proj.bias"])
    x = x * x_mask
    if taps is not None:
        taps["prenet"] = x
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)  # [b,1,t,t]
    K = m["kernel_size"]
    for i in range(m["n_layers_enc"]):
        p = f"encoder.encoder.attn_layers.{i}"
        x = x * x_mask
        q = F.conv1d(x, sd[p + ".conv_q.weight"], sd[p + ".conv_q.bias"])
        k = F.conv1d(x, sd[p + ".conv_k.weight"], sd[p + ".conv_k.bias"])
        v = F.conv1d(x, sd[p + ".conv_v.weight"], sd[p + ".conv_v.bias"])
        a = rel_attention(q, k, v, sd[p + ".emb_rel_k"], sd[p + ".emb_rel_v"], attn_mask.squeeze(1),
                          m["n_heads"], m["window_size"])
        y = F.conv1d(a, sd[p + ".conv_o.weight"], sd[p + ".conv_o.bias"])
        x = channel_layer_norm(x + y, sd[f"encoder.encoder.norm_layers_1.{i}.gamma"], sd[f"encoder.encoder.norm_layers_1.{i}.beta"])
        f = f"encoder.encoder.ffn_layers.{i}"
        y = F.conv1d(x * x_mask, sd[f + ".conv_1.weight"], sd[f + ".conv_1.bias"], padding=K // 2)
        y = torch.relu(y)
        y = F.conv1d(y * x_mask, sd[f + ".conv_2.weight"], sd[f + ".conv_2.bias"], padding=K // 2) * x_mask
        x = channel_layer_norm(x + y, sd[f"encoder.encoder.norm_layers_2.{i}.gamma"], sd[f"encoder.encoder.norm_layers_2.{i}.beta"])
        if taps is not None:
            taps[f"enc{i}"] = x * x_mask
    x = x * x_mask
    x_m = F.conv1d(x, sd["encoder.proj_m.weight"], sd["encoder.proj_m.bias"]) * x_mask
    # duration predictor: conv -> ReLU -> LN (x2) -> 1x1
    h = F.conv1d(x * x_mask, sd["encoder.proj_w.conv_1.weight"], sd["encoder.proj_w.conv_1.bias"], padding=K // 2)
    h = channel_layer_norm(torch.relu(h), sd["encoder.proj_w.norm_1.gamma"], sd["encoder.proj_w.norm_1.beta"])
    h = F.conv1d(h * x_mask, sd["encoder.proj_w.conv_2.weight"], sd["encoder.proj_w.conv_2.bias"], padding=K // 2)
    h = channel_layer_norm(torch.relu(h), sd["encoder.proj_w.norm_2.gamma"], sd["encoder.proj_w.norm_2.beta"])
    logw = F.conv1d(h * x_mask, sd["encoder.proj_w.proj.weight"], sd["encoder.proj_w.proj.bias"]) * x_mask
    return x_m, logw, x_mask


def durations_to_alignment(logw: Tensor, x_mask: Tensor, length_scale: float, n_sqz: int):
    """glow_tts/models.py:323-331,356-363 + utils.py:99-115 as indices.

    Returns w_ceil [b,t_x] (float, integer valued), y_lengths [b] (int64, truncated to a
    multiple of n_sqz), and idx [b, T_y] int64: the phoneme each frame copies (-1 = padding).
    generate_path's one-hot `path[i,t] = (t<cum[i]) - (t<cum[i-1])` is repeat_interleave."""
    w = torch.exp(logw) * x_mask * length_scale
    w_ceil = torch.ceil(w)
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    y_lengths = torch.div(y_lengths, n_sqz, rounding_mode="trunc") * n_sqz
    t_y = int(y_lengths.max())
    cum = torch.cumsum(w_ceil.squeeze(1), 1)  # [b, t_x]
    frames = torch.arange(t_y, dtype=cum.dtype).view(1, 1, t_y)
    idx = (frames >= cum.unsqueeze(-1)).sum(1)  # number of phonemes whose span ended at or before t
    valid = length_mask(y_lengths, t_y)
    idx = torch.where(valid, idx, torch.full_like(idx, -1))
    return w_ceil.squeeze(1), y_lengths, idx


def invconv_near_inverse(x: Tensor, weight: Tensor, x_mask: Tensor, n_split: int = 4) -> Tensor:
    """glow_tts/layers.py:238-272 reverse path with weight_inv = inverse(weight) (layers.py:274-275)."""
    b, c, t = x.shape
    w_inv = torch.inverse(weight.float())
    xg = x.view(b, 2, c // n_split, n_split // 2, t).permute(0, 1, 3, 2, 4).reshape(b, n_split, c // n_split, t)
    z = torch.einsum("og,bgqt->boqt", w_inv, xg)
    z = z.view(b, 2, n_split // 2, c // n_split, t).permute(0, 1, 3, 2, 4).reshape(b, c, t)
    return z * x_mask


def flow_decoder_reverse(sd: SD, cfg: dict, z: Tensor, z_mask: Tensor, taps: Optional[dict] = None) -> Tensor:
    """glow_tts/models.py:191-209 with reverse=True: squeeze, reversed(flows), unsqueeze."""
    m = cfg["model"]
    n_sqz = m["n_sqz"]
    b, c, t = z.shape
    assert t % n_sqz == 0
    # squeeze utils.py:135-147
    x = z.view(b, c, t // n_sqz, n_sqz).permute(0, 3, 1, 2).reshape(b, c * n_sqz, t // n_sqz)
    mask = z_mask[:, :, n_sqz - 1 :: n_sqz]
    x = x * mask
    C = c * n_sqz
    Hd = m["hidden_channels_dec"] or m["hidden_channels"]
    kd = m["kernel_size_dec"]
    nl = m["n_block_layers"]
    for blk in reversed(range(m["n_blocks_dec"])):
        a, inv, cp = 3 * blk, 3 * blk + 1, 3 * blk + 2
        p = f"decoder.flows.{cp}"
        # CouplingBlock reverse attentions.py:119-142
        x0, x1 = x[:, : C // 2], x[:, C // 2 :]
        h = F.conv1d(x0, fold_weight_norm(sd, p + ".start"), sd[p + ".start.bias"]) * mask
        out = torch.zeros_like(h)
        for l in range(nl):  # WN layers.py:138-162
            dil = m["dilation_rate"] ** l
            pad = (kd * dil - dil) // 2
            a_in = F.conv1d(h, fold_weight_norm(sd, p + f".wn.in_layers.{l}"), sd[p + f".wn.in_layers.{l}.bias"],
                            dilation=dil, padding=pad)
            acts = torch.tanh(a_in[:, :Hd]) * torch.sigmoid(a_in[:, Hd:])  # utils.py:31-38
            rs = F.conv1d(acts, fold_weight_norm(sd, p + f".wn.res_skip_layers.{l}"), sd[p + f".wn.res_skip_layers.{l}.bias"])
            if l < nl - 1:
                h = (h + rs[:, :Hd]) * mask
                out = out + rs[:, Hd:]
            else:
                out = out + rs
        out = out * mask
        o = F.conv1d(out, sd[p + ".end.weight"], sd[p + ".end.bias"])
        mu, logs = o[:, : C // 2], o[:, C // 2 :]
        z1 = (x1 - mu) * torch.exp(-logs) * mask
        x = torch.cat([x0, z1], 1)
        # InvConvNear reverse, ActNorm reverse layers.py:182-199
        x = invconv_near_inverse(x, sd[f"decoder.flows.{inv}.weight"], mask, m["n_split"])
        x = (x - sd[f"decoder.flows.{a}.bias"]) * torch.exp(-sd[f"decoder.flows.{a}.logs"]) * mask
        if taps is not None:
            taps[f"flow{blk}"] = x
    # unsqueeze utils.py:150-160
    y = x.view(b, n_sqz, C // n_sqz, t // n_sqz).permute(0, 2, 3, 1).reshape(b, C // n_sqz, t)
    return y * z_mask


def glow_forward(sd: SD, cfg: dict, ids: Tensor, x_lengths: Tensor, noise_scale: float = 0.667,
                 length_scale: float = 1.0, eps: Optional[Tensor] = None, taps: Optional[dict] = None):
    """glow_tts/models.py:308-354 (FlowGenerator.forward, inference branch, g=None).

    `eps` [b, 80, >=T_y] replaces torch.randn_like (models.py:348) so runs are comparable.
    Returns dict(mel [b,80,T_y], y_lengths, w_ceil, idx, logw, x_m, z)."""
    with torch.no_grad():
        x_m, logw, x_mask = text_encoder(sd, cfg, ids, x_lengths, taps)
        w_ceil, y_lengths, idx = durations_to_alignment(logw, x_mask, length_scale, cfg["model"]["n_sqz"])
        t_y = idx.shape[1]
        z_mask = length_mask(y_lengths, t_y).unsqueeze(1).to(x_m.dtype)
        gather = idx.clamp(min=0, max=x_m.shape[2] - 1).unsqueeze(1).expand(-1, x_m.shape[1], -1)
        z_m = torch.gather(x_m, 2, gather) * z_mask  # attn^T . x_m, models.py:336-340
        if eps is None:
            eps = torch.zeros_like(z_m)
        z = (z_m + eps[:, :, :t_y] * noise_scale) * z_mask  # z_logs = 0 (mean_only)
        mel = flow_decoder_reverse(sd, cfg, z, z_mask, taps) if t_y > 0 else z
    return {"mel": mel, "y_lengths": y_lengths, "w_ceil": w_ceil, "idx": idx, "logw": logw, "x_m": x_m, "z": z}


# -----------------------------------------------------------------------------
# host glue between the two passes
# -----------------------------------------------------------------------------


def mel_glue(mel: np.ndarray, audio: dict) -> np.ndarray:
    """larynx/__init__.py:242-249 calling larynx/audio.py:83-104 (denormalize), :58-59 (db_to_amp),
    :106-108 (dynamic_range_compression), in float32 numpy like the reference."""
    m = np.asarray(mel, dtype=np.float32)
    if audio.get("signal_norm", False):
        mx, lo = audio["max_norm"], audio["min_level_db"]
        if audio.get("symmetric_norm", True):
            if audio.get("clip_norm", True):
                m = np.clip(m, -mx, mx)
            m = ((m + mx) * -lo / (2 * mx)) + lo
        else:
            if audio.get("clip_norm", True):
                m = np.clip(m, 0, mx)
            m = (m * -lo / mx) + lo
        m = m + audio["ref_level_db"]
    if audio.get("convert_db_to_amp", True):
        m = np.power(10.0, m / audio["spec_gain"])
    if audio.get("do_dynamic_range_compression", True):
        m = np.log(np.clip(m, a_min=1e-5, a_max=None))
    return m.astype(np.float32)


def float_to_int16(audio: np.ndarray, max_wav_value: float = 32767.0) -> np.ndarray:
    """larynx/audio.py:118-125."""
    a = audio * (max_wav_value / max(0.01, np.max(np.abs(audio)))) if audio.size else audio
    return np.clip(a, -max_wav_value, max_wav_value).astype("int16")


# -----------------------------------------------------------------------------
# HiFi-GAN
# -----------------------------------------------------------------------------


def hifigan_forward(sd: SD, cfg: dict, mel: Tensor, taps: Optional[dict] = None) -> Tensor:
    """hifi_gan/models.py:186-202 (Generator.forward) with ResBlock1 :91-98 / ResBlock2 :137-142.
    mel [b,80,T] -> audio [b,1,T*prod(upsample_rates)] in (-1,1)."""
    m = cfg["model"]
    nk = len(m["resblock_kernel_sizes"])
    with torch.no_grad():
        x = F.conv1d(mel, fold_weight_norm(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
        for i, (u, k) in enumerate(zip(m["upsample_rates"], m["upsample_kernel_sizes"])):
            x = F.leaky_relu(x, 0.1)
            x = F.conv_transpose1d(x, fold_weight_norm(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
            xs = None
            for j, (rk, dils) in enumerate(zip(m["resblock_kernel_sizes"], m["resblock_dilation_sizes"])):
                p = f"resblocks.{i * nk + j}"
                r = x
                for di, d in enumerate(dils):
                    if m["resblock"] == "1":
                        xt = F.leaky_relu(r, 0.1)
                        xt = F.conv1d(xt, fold_weight_norm(sd, p + f".convs1.{di}"), sd[p + f".convs1.{di}.bias"],
                                      dilation=d, padding=(rk * d - d) // 2)
                        xt = F.leaky_relu(xt, 0.1)
                        xt = F.conv1d(xt, fold_weight_norm(sd, p + f".convs2.{di}"), sd[p + f".convs2.{di}.bias"],
                                      padding=(rk - 1) // 2)
                    else:
                        xt = F.leaky_relu(r, 0.1)
                        xt = F.conv1d(xt, fold_weight_norm(sd, p + f".convs.{di}"), sd[p + f".convs.{di}.bias"],
                                      dilation=d, padding=(rk * d - d) // 2)
                    r = xt + r
                xs = r if xs is None else xs + r
            x = xs / nk
            if taps is not None:
                taps[f"stage{i}"] = x
        x = F.leaky_relu(x)  # default slope 0.01, hifi_gan/models.py:198
        x = F.conv1d(x, fold_weight_norm(sd, "conv_post"), sd["conv_post.bias"], padding=3)
        return torch.tanh(x)


# -----------------------------------------------------------------------------
# denoiser (host numpy in the reference too)
# -----------------------------------------------------------------------------


def stft_frames(x: np.ndarray, fft_size: int = 1024, hop: int = 256) -> np.ndarray:
    """larynx/audio.py:232-248 — Hann window, frames range(0, len(x) - fft_size, hop); rows = time slices."""
    window = np.hanning(fft_size)
    return np.array([np.fft.rfft(window * x[i : i + fft_size]) for i in range(0, len(x) - fft_size, hop)])


def denoise_reference(audio: np.ndarray, bias_spec: np.ndarray, strength: float) -> np.ndarray:
    """larynx/hifi_gan.py:171-179 with transform (audio.py:291-306) and inverse/istft (audio.py:251-288).
    audio [1, N] -> [1, n_t*256 + 1024] float64."""
    spec = stft_frames(audio[0]).T
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2)[None]
    phase = np.arctan2(spec.imag, spec.real)[None]
    mag = np.clip(mag - bias_spec * strength, a_min=0.0, a_max=None)
    n_f = mag.shape[1]
    comb = np.concatenate([mag * np.cos(phase), mag * np.sin(phase)], axis=1)
    x = np.empty([1, n_f, comb.shape[2]], dtype=np.complex64)
    x.real = comb[:, :n_f]
    x.imag = comb[:, n_f:]
    X = x[0].T
    window = np.hanning(1024)
    out = np.zeros(int(X.shape[0] * 256 + 1024))
    for n, i in enumerate(range(0, len(out) - 1024, 256)):
        out[i : i + 1024] += window * np.real(np.fft.irfft(X[n]))
    return out[None, :]


def denoiser_bias_spec(hifi_sd: SD, hifi_cfg: dict) -> np.ndarray:
    """larynx/hifi_gan.py:181-203 — first STFT frame of the vocoder run on zeros [1, 80, 88]."""
    bias_audio = hifigan_forward(hifi_sd, hifi_cfg, torch.zeros(1, 80, 88)).squeeze(0).numpy()
    spec = stft_frames(bias_audio[0]).T
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2)[None]
    return mag[:, :, 0][:, :, None]


# -----------------------------------------------------------------------------
# whole path (== larynx/__init__.py:214-285 _sentence_task without logging / SSML pauses)
# -----------------------------------------------------------------------------


def sentence_task(glow_sd: SD, glow_cfg: dict, hifi_sd: SD, hifi_cfg: dict, ids: np.ndarray,
                  noise_scale: float = 0.667, length_scale: float = 1.0, eps: Optional[np.ndarray] = None):
    ids_t = torch.from_numpy(np.asarray(ids, dtype=np.int64)).unsqueeze(0)
    lens = torch.tensor([ids_t.shape[1]], dtype=torch.int64)
    eps_t = None if eps is None else torch.from_numpy(eps).unsqueeze(0)
    g = glow_forward(glow_sd, glow_cfg, ids_t, lens, noise_scale, length_scale, eps_t)
    mel = g["mel"].numpy()
    mel2 = mel_glue(mel, glow_cfg["audio"])
    pcm = hifigan_forward(hifi_sd, hifi_cfg, torch.from_numpy(mel2)).squeeze(0).numpy()
    return {"mel": mel, "mel_glued": mel2, "pcm": pcm.squeeze(0), "audio": float_to_int16(pcm).squeeze(),
            "y_lengths": g["y_lengths"].numpy(), "w_ceil": g["w_ceil"].numpy(), "idx": g["idx"].numpy(),
            "logw": g["logw"].numpy()}
