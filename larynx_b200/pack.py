"""Weight packer: reference state-dicts -> one contiguous blob per model, in kernel layout.

Load-time work of the reference that is folded in here (so the kernels never see it):
  * weight-norm folding w = g * v / ||v|| — `CouplingBlock.start` keeps weight_g/weight_v at run
    time (glow_tts/attentions.py:100-102), `WN.remove_weight_norm` (glow_tts/layers.py:164-170)
    and `Generator.remove_weight_norm` (hifi_gan/models.py:204-211) fold at load;
  * `InvConvNear.store_inverse` (glow_tts/layers.py:274-275): W^-1 in fp32;
  * ActNorm's exp(-logs) (glow_tts/layers.py:195).
Kernel layouts:
  * Conv1d weight [Cout][Cin][K] -> [Cin][K][Cout] (Cout contiguous: vector loads per thread);
  * WN in_layer / coupling `end`: output channels interleaved (c, c+Cout/2) -> (2c, 2c+1) so the
    tanh*sigmoid gate / (m, logs) pair sits in one thread;
  * q/k/v 1x1 convs concatenated into one [H][1][3H] conv;
  * ConvTranspose1d [Cin][Cout][2u] -> u polyphase 2-tap convs [u][Cin][2][Cout].
Blob = header | tensor table | 256-byte aligned fp32 payloads (csrc/engine.cu:BlobHeader/BlobEntry).
This blob is also the unit that is broadcast once over NCCL when several GPUs serve one voice.
"""
from __future__ import annotations

import struct
from typing import Optional, Dict, List, Tuple

import json

import numpy as np

MAGIC = b"B2TTSBLB"
KIND_GLOW = 1
KIND_HIFIGAN = 2
_HEADER = struct.Struct("<8sIIII64iQ")
_ENTRY = struct.Struct("<64sQQII")


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def fold_weight_norm(sd: Dict[str, np.ndarray], prefix: str) -> np.ndarray:
    """w = g * v / ||v||_2 with the norm over all dims but 0 (torch.nn.utils.weight_norm, dim=0)."""
    if prefix + ".weight" in sd:
        return _np(sd[prefix + ".weight"])
    v = _np(sd[prefix + ".weight_v"])
    g = _np(sd[prefix + ".weight_g"])
    norm = np.sqrt(np.sum(v.astype(np.float32) ** 2, axis=tuple(range(1, v.ndim)), keepdims=True, dtype=np.float32))
    return (v * (g.reshape(norm.shape) / norm)).astype(np.float32)


def conv_layout(w: np.ndarray) -> np.ndarray:
    """[Cout][Cin][K] -> [Cin][K][Cout]."""
    return np.ascontiguousarray(np.transpose(w, (1, 2, 0)))


def interleave_halves(w: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Reorder output channels (c, c+C/2) -> (2c, 2c+1).  w is [Cout][Cin][K]."""
    co = w.shape[0]
    order = np.empty(co, dtype=np.int64)
    order[0::2] = np.arange(co // 2)
    order[1::2] = np.arange(co // 2) + co // 2
    return w[order], b[order]


def polyphase(w: np.ndarray, u: int, k: int) -> np.ndarray:
    """ConvTranspose1d weight [Cin][Cout][k], stride u, padding (k-u)/2, k == 2u
    -> [u][Cin][2][Cout]: output n = t*u + r uses tap0 on x[t+q-1] and tap1 on x[t+q],
    q = (r+pad)//u, with tap0 = W[..., j0+u], tap1 = W[..., j0], j0 = (r+pad) % u
    (y[n] = sum_i x[i] W[n + pad - i*u], torch ConvTranspose1d)."""
    assert k == 2 * u
    pad = (k - u) // 2
    cin, cout, _ = w.shape
    out = np.empty((u, cin, 2, cout), dtype=np.float32)
    for r in range(u):
        j0 = (r + pad) % u
        out[r, :, 0, :] = w[:, :, j0 + u]
        out[r, :, 1, :] = w[:, :, j0]
    return out


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit patterns (uint16), round to nearest even."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    bias = np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    return ((u + bias) >> np.uint32(16)).astype(np.uint16)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def pick_nt(cout: int, nt_max: int = 128) -> int:
    """Output-channel tile of the tensor-core conv: a multiple of 16, <= nt_max, dividing Cout rounded up to 16.
    Layers that run on tiny grids (GlowTTS at batch 1: one or two 128-row time tiles) use a smaller tile so more
    SMs share the serial per-tile latency."""
    c16 = (cout + 15) // 16 * 16
    for nt in (128, 112, 96, 80, 64, 48, 32, 16):
        if nt <= nt_max and c16 % nt == 0:
            return nt
    return 16


UMMA_BKC = 32  # input channels per pipeline stage (csrc/conv_umma.cuh: UM_BKC)


def umma_image(w: np.ndarray, nsplit: int, nt_max: int = 128):
    """[Cout][Cin][K] fp32 -> bf16 split image for conv_umma_kernel:
    [n-tile][ci-chunk of 32][tap][k-chunk of 8 (4)][split plane hi/lo(/lo2)][NT rows][8] uint16.
    Each (n-tile, ci-chunk, tap) block is one contiguous cp.async.bulk transfer and already has the
    K-major SWIZZLE_NONE UMMA operand layout (rows 16 B apart).  The split planes of one k-chunk are adjacent
    row blocks, so ONE MMA with N = 2*NT (3*NT) multiplies an activation plane with hi|lo(|lo2) at once."""
    cout, cin, k = w.shape
    nt = pick_nt(cout, nt_max)
    ntiles = (cout + nt - 1) // nt
    nch = (cin + UMMA_BKC - 1) // UMMA_BKC
    wp = np.zeros((ntiles * nt, nch * UMMA_BKC, k), dtype=np.float32)
    wp[:cout, :cin] = w
    planes = []
    rem = wp
    for _ in range(nsplit):
        bits = to_bf16_bits(rem)
        planes.append(bits)
        rem = (rem - from_bf16_bits(bits)).astype(np.float32)
    img = np.stack(planes, 0)  # [split][N][C][K]
    img = img.reshape(nsplit, ntiles, nt, nch, UMMA_BKC // 8, 8, k)
    img = np.transpose(img, (1, 3, 6, 4, 0, 2, 5))  # [ntile][nch][K][kc][split][NT][8]
    meta = np.array([nt, nsplit, nch, k, ntiles, cout, cin, 0], dtype=np.float32)
    return np.ascontiguousarray(img).reshape(-1), meta


def rel_image(e: np.ndarray) -> np.ndarray:
    """Relative-position embedding [2w+1][96] -> bf16 3-way split operand image [plane][12][16][8] (rows >= 2w+1 zero)."""
    n, d = e.shape
    full = np.zeros((16, d), dtype=np.float32)
    full[:n] = e
    planes = []
    rem = full
    for _ in range(3):
        bits = to_bf16_bits(rem)
        planes.append(bits)
        rem = (rem - from_bf16_bits(bits)).astype(np.float32)
    img = np.stack(planes, 0).reshape(3, 16, d // 8, 8)  # [plane][row][group][8]
    return np.ascontiguousarray(np.transpose(img, (0, 2, 1, 3))).reshape(-1)


def pixel_shuffle_conv(w: np.ndarray, u: int, k: int) -> np.ndarray:
    """ConvTranspose1d weight [Cin][Cout][k] (stride u, padding (k-u)/2, k == 2u) as ONE ordinary 3-tap conv
    (pad_left 1) with u*Cout packed output channels n = co*u + r:  y[co][t*u + r] = conv3(x)[n][t].
    Phase r reads x[t+q-1] with W[.., j0+u] and x[t+q] with W[.., j0] (q, j0 as in polyphase()); the third
    tap of every phase is zero."""
    assert k == 2 * u
    pad = (k - u) // 2
    cin, cout, _ = w.shape
    out = np.zeros((cout * u, cin, 3), dtype=np.float32)
    for r in range(u):
        q, j0 = divmod(r + pad, u)
        out[r::u, :, q] = w[:, :, j0 + u].T
        out[r::u, :, q + 1] = w[:, :, j0].T
    return out


class BlobWriter:
    def __init__(self, kind: int, cfg: List[int]):
        self.kind = kind
        self.cfg = list(cfg)
        self.items: List[Tuple[str, np.ndarray]] = []

    def add(self, name: str, arr: np.ndarray):
        assert len(name) < 64, name
        self.items.append((name, np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)))

    def add_u16(self, name: str, bits: np.ndarray):
        """uint16 payload stored as a float32-typed entry (numel counts 4-byte words)."""
        bits = np.ascontiguousarray(bits, dtype=np.uint16).reshape(-1)
        if bits.size % 2:
            bits = np.concatenate([bits, np.zeros(1, np.uint16)])
        self.items.append((name, bits.view(np.float32)))

    nt_max = 128
    simt = False  # also store the fp32 [Cin][K][Cout] copies the SIMT cross-check path reads (B200TTS_DEBUG=simt)

    def add_conv(self, name: str, w: np.ndarray, b: np.ndarray, nsplit: int = 2, simt: Optional[bool] = None,
                 nt_max: Optional[int] = None):
        """w is [Cout][Cin][K] in final (packed) channel order."""
        if self.simt if simt is None else simt:
            self.add(name + ".w", conv_layout(w))
        self.add(name + ".b", b)
        img, meta = umma_image(w, nsplit, self.nt_max if nt_max is None else nt_max)
        self.add_u16(name + ".u", img)
        self.add(name + ".um", meta)

    def finish(self) -> np.ndarray:
        n = len(self.items)
        off = _HEADER.size + _ENTRY.size * n
        off = (off + 255) // 256 * 256
        entries = []
        for name, arr in self.items:
            entries.append((name, off, arr.size))
            off += (arr.size * 4 + 255) // 256 * 256
        total = off
        blob = np.zeros(total, dtype=np.uint8)
        cfg = self.cfg + [0] * (64 - len(self.cfg))
        blob[: _HEADER.size] = np.frombuffer(_HEADER.pack(MAGIC, 1, self.kind, n, len(self.cfg), *cfg, total), dtype=np.uint8)
        p = _HEADER.size
        for (name, o, numel), (_, arr) in zip(entries, self.items):
            blob[p : p + _ENTRY.size] = np.frombuffer(_ENTRY.pack(name.encode(), o, numel, 0, 0), dtype=np.uint8)
            p += _ENTRY.size
            blob[o : o + numel * 4] = arr.view(np.uint8)
        return blob


def want_simt_weights() -> bool:
    """The fp32 SIMT conv path is an in-library cross-check (B200TTS_DEBUG=simt); only then do the blobs carry its weights
    (they double the blob: 243 MB instead of 126 MB per GlowTTS voice)."""
    import os

    return "simt" in os.environ.get("B200TTS_DEBUG", "").split(",")


def pack_glow(sd: Dict[str, np.ndarray], config: dict, simt: Optional[bool] = None) -> np.ndarray:
    """Pack a glow_tts FlowGenerator state-dict (reference key names) for b200tts_glow_create."""
    m = config["model"]
    if m.get("n_speakers", 1) > 1 or m.get("gin_channels", 0) != 0:
        raise ValueError("multi-speaker GlowTTS voices are not supported")
    if not m.get("mean_only", True) or not m.get("prenet", True) or m.get("sigmoid_scale", False):
        raise ValueError("unsupported GlowTTS variant (needs mean_only, prenet, no sigmoid_scale)")
    if m.get("block_length") is not None:
        raise ValueError("block_length attention is not supported")
    H = m.get("hidden_channels_enc") or m["hidden_channels"]
    Hd = m.get("hidden_channels_dec") or m["hidden_channels"]
    mel = config["audio"]["mel_channels"]
    cfg = [m["num_symbols"], H, m["filter_channels"], m["filter_channels_dp"], m["kernel_size"], m["n_heads"],
           m["n_layers_enc"], m["window_size"], mel, m["n_blocks_dec"], m["kernel_size_dec"], m["dilation_rate"],
           m["n_block_layers"], m["n_sqz"], m["n_split"], Hd]
    bw = BlobWriter(KIND_GLOW, cfg)
    bw.simt = want_simt_weights() if simt is None else bool(simt)

    bw.nt_max = 64  # GlowTTS runs on 1-4 time tiles: favour more N tiles (flow_chain.cuh is built for this tile)
    g = lambda k: _np(sd[k])
    bw.add("emb", g("encoder.emb.weight"))
    for i in range(3):
        bw.add_conv(f"pre.conv{i}", g(f"encoder.pre.conv_layers.{i}.weight"), g(f"encoder.pre.conv_layers.{i}.bias"), nsplit=3)
        bw.add(f"pre.ln{i}.g", g(f"encoder.pre.norm_layers.{i}.gamma"))
        bw.add(f"pre.ln{i}.b", g(f"encoder.pre.norm_layers.{i}.beta"))
    bw.add_conv("pre.proj", g("encoder.pre.proj.weight"), g("encoder.pre.proj.bias"), nsplit=3)
    for i in range(m["n_layers_enc"]):
        p = f"encoder.encoder.attn_layers.{i}"
        wq, wk, wv = g(p + ".conv_q.weight"), g(p + ".conv_k.weight"), g(p + ".conv_v.weight")
        bq, bk, bv = g(p + ".conv_q.bias"), g(p + ".conv_k.bias"), g(p + ".conv_v.bias")
        bw.add_conv(f"enc{i}.qkv", np.concatenate([wq, wk, wv], 0), np.concatenate([bq, bk, bv], 0), nsplit=3)
        bw.add_conv(f"enc{i}.o", g(p + ".conv_o.weight"), g(p + ".conv_o.bias"), nsplit=3)
        ek, ev = g(p + ".emb_rel_k"), g(p + ".emb_rel_v")
        if ek.shape[0] != 1:
            raise ValueError("heads_share=False relative embeddings are not supported")
        bw.add(f"enc{i}.rel_k", ek[0])
        bw.add(f"enc{i}.rel_v", ev[0])
        if ek.shape[2] == 96 and ek.shape[1] <= 9:
            # operand images for the tcgen05 attention: rows r = E[r], [3 planes][12 groups][16 rows][8 channels]
            bw.add_u16(f"enc{i}.rel_k_img", rel_image(ek[0]))
            bw.add_u16(f"enc{i}.rel_v_img", rel_image(ev[0]))
        bw.add(f"enc{i}.ln1.g", g(f"encoder.encoder.norm_layers_1.{i}.gamma"))
        bw.add(f"enc{i}.ln1.b", g(f"encoder.encoder.norm_layers_1.{i}.beta"))
        f = f"encoder.encoder.ffn_layers.{i}"
        bw.add_conv(f"enc{i}.ffn1", g(f + ".conv_1.weight"), g(f + ".conv_1.bias"), nsplit=3)
        bw.add_conv(f"enc{i}.ffn2", g(f + ".conv_2.weight"), g(f + ".conv_2.bias"), nsplit=3)
        bw.add(f"enc{i}.ln2.g", g(f"encoder.encoder.norm_layers_2.{i}.gamma"))
        bw.add(f"enc{i}.ln2.b", g(f"encoder.encoder.norm_layers_2.{i}.beta"))
    bw.add_conv("proj_m", g("encoder.proj_m.weight"), g("encoder.proj_m.bias"), nsplit=3)
    bw.add_conv("dp.conv1", g("encoder.proj_w.conv_1.weight"), g("encoder.proj_w.conv_1.bias"), nsplit=3)
    bw.add("dp.ln1.g", g("encoder.proj_w.norm_1.gamma"))
    bw.add("dp.ln1.b", g("encoder.proj_w.norm_1.beta"))
    bw.add_conv("dp.conv2", g("encoder.proj_w.conv_2.weight"), g("encoder.proj_w.conv_2.bias"), nsplit=3)
    bw.add("dp.ln2.g", g("encoder.proj_w.norm_2.gamma"))
    bw.add("dp.ln2.b", g("encoder.proj_w.norm_2.beta"))
    bw.add("dp.proj.w", g("encoder.proj_w.proj.weight"))
    bw.add("dp.proj.b", g("encoder.proj_w.proj.bias"))
    nl = m["n_block_layers"]
    for b in range(m["n_blocks_dec"]):
        a, inv, cp = 3 * b, 3 * b + 1, 3 * b + 2
        p = f"decoder.flows.{cp}"
        bw.add_conv(f"flow{b}.start", fold_weight_norm(sd, p + ".start"), g(p + ".start.bias"))
        for l in range(nl):
            w, bias = interleave_halves(fold_weight_norm(sd, p + f".wn.in_layers.{l}"), g(p + f".wn.in_layers.{l}.bias"))
            bw.add_conv(f"flow{b}.in{l}", w, bias)
            bw.add_conv(f"flow{b}.rs{l}", fold_weight_norm(sd, p + f".wn.res_skip_layers.{l}"),
                        g(p + f".wn.res_skip_layers.{l}.bias"))
        w, bias = interleave_halves(g(p + ".end.weight"), g(p + ".end.bias"))
        bw.add_conv(f"flow{b}.end", w, bias)
        if f"decoder.flows.{inv}.weight_inv" in sd:  # an ONNX export carries the matrix of store_inverse() itself
            winv = g(f"decoder.flows.{inv}.weight_inv").reshape(m["n_split"], m["n_split"])
        else:
            winv = np.linalg.inv(g(f"decoder.flows.{inv}.weight").astype(np.float32)).astype(np.float32)
        bw.add(f"flow{b}.inv", winv)
        bw.add(f"flow{b}.an_bias", g(f"decoder.flows.{a}.bias").reshape(-1))
        bw.add(f"flow{b}.an_scale", np.exp(-g(f"decoder.flows.{a}.logs").reshape(-1)).astype(np.float32))
    return bw.finish()


def pack_hifigan(sd: Dict[str, np.ndarray], config: dict, simt: Optional[bool] = None) -> np.ndarray:
    """Pack a hifi_gan Generator state-dict (reference key names) for b200tts_hifigan_create.
    `config` may be nested ({"model": {...}}) or the flat jik876 json (universal_large/config.json)."""
    m = config.get("model", config)
    rates = list(m["upsample_rates"])
    ks = list(m["upsample_kernel_sizes"])
    rk = list(m["resblock_kernel_sizes"])
    rd = [list(d) for d in m["resblock_dilation_sizes"]]
    C0 = int(m["upsample_initial_channel"])
    resblock = int(m.get("resblock", "1"))
    cfg = [0] * 40
    cfg[0], cfg[1], cfg[2], cfg[3], cfg[4] = resblock, len(rates), C0, len(rk), 80
    for i, (u, k) in enumerate(zip(rates, ks)):
        cfg[8 + i], cfg[16 + i] = u, k
    for j, k in enumerate(rk):
        cfg[24 + j] = k
        for d, dv in enumerate(rd[j]):
            cfg[28 + j * 3 + d] = dv
    bw = BlobWriter(KIND_HIFIGAN, cfg)
    bw.simt = want_simt_weights() if simt is None else bool(simt)
    g = lambda k: _np(sd[k])
    rb_nt = 128  # ResBlock output-channel tile (64-wide tiles measured equal within 1 % on B200: fewer CTAs win)
    bw.add_conv("conv_pre", fold_weight_norm(sd, "conv_pre"), g("conv_pre.bias"))
    for i, (u, k) in enumerate(zip(rates, ks)):
        wt = fold_weight_norm(sd, f"ups.{i}")
        if bw.simt:
            bw.add(f"ups{i}.w", polyphase(wt, u, k))
        bw.add(f"ups{i}.b", g(f"ups.{i}.bias"))
        # tensor-core form: one 3-tap conv with Cout*u pixel-shuffled output channels
        bw.add_conv(f"ups{i}.ps", pixel_shuffle_conv(wt, u, k), np.repeat(g(f"ups.{i}.bias"), u), simt=False)
    n = 0
    for i in range(len(rates)):
        for j in range(len(rk)):
            p = f"resblocks.{n}"
            for d in range(len(rd[j])):
                if resblock == 1:
                    bw.add_conv(f"rb{n}.c1_{d}", fold_weight_norm(sd, p + f".convs1.{d}"), g(p + f".convs1.{d}.bias"), nt_max=rb_nt)
                    bw.add_conv(f"rb{n}.c2_{d}", fold_weight_norm(sd, p + f".convs2.{d}"), g(p + f".convs2.{d}.bias"), nt_max=rb_nt)
                else:
                    bw.add_conv(f"rb{n}.c1_{d}", fold_weight_norm(sd, p + f".convs.{d}"), g(p + f".convs.{d}.bias"), nt_max=rb_nt)
            n += 1
    bw.add("conv_post.w", fold_weight_norm(sd, "conv_post"))
    bw.add("conv_post.b", g("conv_post.bias"))
    return bw.finish()


def load_state_dict(path, key: str) -> Dict[str, np.ndarray]:
    """Read the reference checkpoint format: torch.load(path)[key] with key = "model"
    (glow_tts/checkpoint.py:33-41) or "generator" (hifi_gan/checkpoint.py:47-49)."""
    import torch

    ckpt = torch.load(str(path), map_location="cpu", weights_only=False)
    if key not in ckpt:
        raise KeyError(f"Missing '{key}' in state dict")
    sd = ckpt[key]
    return {(k[7:] if k.startswith("module.") else k): v.detach().cpu().float().numpy() for k, v in sd.items()}


def load_voice_weights(model_dir, key: str) -> Tuple[Dict[str, np.ndarray], str]:
    """The weights of a voice / vocoder directory: `generator.pth` (the reference's PyTorch backend, larynx/glow_tts.py:66-84,
    larynx/hifi_gan.py:70-99) when it exists, else the initializers of `generator.onnx` (what released voices ship) read by
    name (onnx_weights.py).  Returns (state dict, the file it came from)."""
    from pathlib import Path

    d = Path(model_dir)
    pth, onx = d / "generator.pth", d / "generator.onnx"
    if pth.is_file():
        return load_state_dict(pth, key), str(pth)
    if onx.is_file():
        from . import onnx_weights

        if key == "model":  # GlowTTS: some parameters only survive as anonymous constants (onnx_weights.glow_state_dict)
            with open(d / "config.json", "r", encoding="utf-8") as f:
                return onnx_weights.glow_state_dict(onx, json.load(f)), str(onx)
        return onnx_weights.state_dict(onx), str(onx)
    raise FileNotFoundError(f"neither {pth} nor {onx} exists")


def pack_from_directory(model_dir, key: str, config: dict, packer) -> np.ndarray:
    """packer(state dict, config) on the directory's weights; a graph whose export folded parameters into anonymous
    constants (so that names the packer needs are gone) fails with those names."""
    sd, src = load_voice_weights(model_dir, key)
    try:
        return packer(sd, config)
    except KeyError as e:
        if src.endswith(".onnx"):
            raise KeyError(f"{src}: no initializer named {e} — this export folded the parameter into an anonymous constant; "
                           "larynx_b200 reads ONNX weights by parameter name only (use the voice's generator.pth)") from e
        raise
