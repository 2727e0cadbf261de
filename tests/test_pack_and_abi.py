"""Host-side logic: weight packer layouts, blob format, and the C-ABI surface (no GPU needed)."""
import ctypes
import os
import re

import pytest
import numpy as np
import torch
import torch.nn.functional as F

from larynx_b200 import _lib, pack
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fold_weight_norm_matches_torch():
    torch.manual_seed(0)
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(6, 10, 5))
    conv.weight_g.data.mul_(torch.rand(10, 1, 1) + 0.5)
    sd = {k: v.detach().numpy() for k, v in conv.state_dict().items()}
    sd = {"c." + k: v for k, v in sd.items()}
    w = pack.fold_weight_norm(sd, "c")
    x = torch.randn(1, 6, 20)
    assert np.abs(conv(x).detach().numpy() - F.conv1d(x, torch.from_numpy(w), conv.bias).detach().numpy()).max() < 1e-5
    # ConvTranspose1d: weight_norm dim=0 is the *input* channel axis
    ct = torch.nn.utils.weight_norm(torch.nn.ConvTranspose1d(6, 4, 8, 4, padding=2))
    sd = {"t." + k: v.detach().numpy() for k, v in ct.state_dict().items()}
    assert sd["t.weight_g"].shape == (6, 1, 1)
    w = pack.fold_weight_norm(sd, "t")
    assert np.abs(ct(x).detach().numpy() - F.conv_transpose1d(x, torch.from_numpy(w), ct.bias, stride=4, padding=2).detach().numpy()).max() < 1e-5


def test_polyphase_equals_conv_transpose():
    rs = np.random.RandomState(0)
    for u in (8, 2, 4):
        k = 2 * u
        cin, cout, L = 5, 3, 17
        w = rs.standard_normal((cin, cout, k)).astype(np.float32)
        x = rs.standard_normal((cin, L)).astype(np.float32)
        ref = F.conv_transpose1d(torch.from_numpy(x)[None], torch.from_numpy(w), stride=u, padding=(k - u) // 2)[0].numpy()
        pw = pack.polyphase(w, u, k)  # [u][cin][2][cout]
        pad = (k - u) // 2
        out = np.zeros((cout, L * u), dtype=np.float64)
        xp = np.pad(x, ((0, 0), (2, 2)))
        for r in range(u):
            q = (r + pad) // u
            for t in range(L):
                a = xp[:, t + q - 1 + 2] @ pw[r, :, 0, :] + xp[:, t + q + 2] @ pw[r, :, 1, :]
                out[:, t * u + r] = a
        assert ref.shape == out.shape
        assert np.abs(ref - out).max() < 1e-4


def test_interleave_and_conv_layout():
    w = np.arange(8 * 3 * 2, dtype=np.float32).reshape(8, 3, 2)
    b = np.arange(8, dtype=np.float32)
    wi, bi = pack.interleave_halves(w, b)
    assert list(bi) == [0, 4, 1, 5, 2, 6, 3, 7]
    assert np.array_equal(wi[1], w[4])
    cl = pack.conv_layout(w)
    assert cl.shape == (3, 2, 8) and cl[1, 0, 5] == w[5, 1, 0]


def test_blob_roundtrip_table():
    cfg = common.glow_cfg()
    blob = pack.pack_glow(common.glow_sd_np(), cfg)
    magic, ver, kind, n, ncfg, *rest = pack._HEADER.unpack_from(blob.tobytes()[: pack._HEADER.size])
    assert magic == pack.MAGIC and ver == 1 and kind == pack.KIND_GLOW and ncfg == 16
    assert rest[-1] == blob.nbytes
    assert rest[0] == 46 and rest[1] == 192 and rest[8] == 80
    names = {}
    p = pack._HEADER.size
    raw = blob.tobytes()
    for _ in range(n):
        name, off, numel, dt, _r = pack._ENTRY.unpack_from(raw, p)
        names[name.rstrip(b"\0").decode()] = (off, numel)
        assert off % 256 == 0 and off + numel * 4 <= blob.nbytes
        p += pack._ENTRY.size
    off, numel = names["flow3.an_scale"]
    got = blob[off : off + numel * 4].view(np.float32)
    assert np.allclose(got, np.exp(-common.glow_sd_np()["decoder.flows.9.logs"].reshape(-1)))
    off, numel = names["flow0.inv"]
    winv = blob[off : off + 64].view(np.float32).reshape(4, 4)
    assert np.abs(winv @ common.glow_sd_np()["decoder.flows.1.weight"] - np.eye(4)).max() < 1e-5
    # 28.57 M parameters: the default blob holds the bf16 split images of the tensor-core path (2 planes for
    # the flow decoder, 3 for the text encoder) and no fp32 conv copies; the SIMT cross-check path asks for them explicitly
    assert not any(k.endswith(".w") and not k.startswith("dp.proj") for k in names), "fp32 conv copies in the default blob"
    assert 110e6 < blob.nbytes < 160e6, blob.nbytes
    both = pack.pack_glow(common.glow_sd_np(), cfg, simt=True)
    assert 220e6 < both.nbytes < 320e6, both.nbytes


def test_synth_param_counts():
    # GlowTTS 28,569,745 parameters counting weight_v only for the weight-normed convs
    n_glow = sum(v.size for k, v in common.glow_sd_np().items() if not k.endswith("weight_g"))
    assert abs(n_glow - 28_569_745) < 4096
    sd = common.hifi_sd_np()
    n_folded = sum(v.size for k, v in sd.items() if not k.endswith("weight_g"))
    assert n_folded == 13_926_017


def test_library_exports_every_declared_symbol(built_library):
    header = open(os.path.join(ROOT, "include", "b200tts.h")).read()
    declared = set(re.findall(r"\b(b200tts_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200tts_call_set_debug(call,1)"}
    assert len(declared) >= 20
    lib = built_library
    for name in declared:
        assert hasattr(lib, name), f"libb200tts.so does not export {name}"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.b200tts_version() == 1


def test_fails_loudly_without_device_or_with_bad_blob(built_library):
    lib = built_library
    try:
        ngpu = torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        ngpu = 0
    h = ctypes.c_void_p()
    junk = np.zeros(4096, dtype=np.uint8)
    rc = lib.b200tts_glow_create(junk.ctypes.data, junk.nbytes, 0, ctypes.byref(h))
    assert rc != 0
    msg = lib.b200tts_last_error().decode()
    assert msg  # either "bad blob magic" (GPU present) or a CUDA error (no device): never a silent fallback
    if ngpu == 0:
        assert "CUDA" in msg


def test_bf16_split_images_reconstruct_weights():
    """umma_image: [n-tile][ci-chunk][tap][4][plane][NT][8]; hi + lo (+ lo2) must reconstruct the fp32 weights."""
    rs = np.random.RandomState(2)
    for cout, cin, k, nsplit, nt_max in ((160, 192, 1, 2, 64), (384, 192, 5, 2, 128), (80, 192, 1, 3, 64), (64, 64, 11, 2, 128)):
        w = rs.standard_normal((cout, cin, k)).astype(np.float32)
        img, meta = pack.umma_image(w, nsplit, nt_max)
        nt, ns, nch, kk, ntiles = (int(meta[i]) for i in range(5))
        assert nt == pack.pick_nt(cout, nt_max) and ns == nsplit and kk == k and nch == (cin + 31) // 32
        assert nt % 16 == 0 and nt <= nt_max and ntiles * nt >= cout
        planes = pack.from_bf16_bits(img.reshape(ntiles, nch, k, 4, nsplit, nt, 8))
        rec = planes.sum(axis=4)  # [ntile][nch][k][4][nt][8]
        rec = np.transpose(rec, (0, 4, 1, 3, 5, 2)).reshape(ntiles * nt, nch * 32, k)[:cout, :cin]
        tol = 2.0 ** (-16 if nsplit == 2 else -23) * np.abs(w).max()
        assert np.abs(rec - w).max() <= tol
        # padding (channels beyond Cout / Cin) is exactly zero, so garbage x 0 never reaches an accumulator
        full = np.transpose(planes.sum(axis=4), (0, 4, 1, 3, 5, 2)).reshape(ntiles * nt, nch * 32, k)
        assert np.all(full[cout:] == 0) and np.all(full[:, cin:] == 0)


def test_relative_embedding_image():
    e = np.random.RandomState(4).standard_normal((9, 96)).astype(np.float32)
    img = pack.from_bf16_bits(pack.rel_image(e).reshape(3, 12, 16, 8))
    rec = np.transpose(img.sum(axis=0), (1, 0, 2)).reshape(16, 96)
    assert np.abs(rec[:9] - e).max() <= 2.0 ** -22 * np.abs(e).max()
    assert np.all(rec[9:] == 0)


def test_pixel_shuffle_conv_equals_conv_transpose():
    rs = np.random.RandomState(6)
    for u in (8, 2, 4):
        k, cin, cout, L = 2 * u, 6, 4, 19
        w = rs.standard_normal((cin, cout, k)).astype(np.float32)
        x = rs.standard_normal((1, cin, L)).astype(np.float32)
        ref = F.conv_transpose1d(torch.from_numpy(x), torch.from_numpy(w), stride=u, padding=(k - u) // 2)[0].numpy()
        y = F.conv1d(torch.from_numpy(x), torch.from_numpy(pack.pixel_shuffle_conv(w, u, k)), padding=1)[0].numpy()
        out = y.reshape(cout, u, L).transpose(0, 2, 1).reshape(cout, L * u)  # packed channel n = co*u + r
        assert np.abs(out - ref).max() < 1e-5


def test_stage_partition_is_contiguous_and_balanced(built_library):
    """Host logic of the per-stage HiFi-GAN kernels (stage_umma.cu: stage_partition): every tile goes to exactly one CTA, CTA
    indices never decrease along the strips (contiguous ranges), and every CTA's modelled cost is within one tile of the mean —
    for the bench shapes (3 branches x 251 tiles, costs of the 32-channel stage), a ragged batch and more CTAs than tiles."""
    import ctypes as C

    import numpy as np

    from larynx_b200 import _lib

    lib = _lib.load()

    def run(tiles, cost, nctas):
        tiles = np.asarray(tiles, np.int32)
        cost = np.asarray(cost, np.int64)
        out = np.full(int(tiles.sum()), -7, np.int32)
        nr = C.c_int32(0)
        rc = lib.b200tts_debug_stage_partition(tiles.ctypes.data, cost.ctypes.data, len(tiles), nctas, out.ctypes.data, C.byref(nr))
        assert rc == 0, lib.b200tts_last_error()
        return out, int(nr.value)

    for tiles, cost, nctas in (([251, 251, 251], [3000, 3000, 3594], 148),
                               ([5, 34, 12, 5, 34, 12, 5, 34, 12], [3600, 3600, 3600, 5088, 5088, 5088, 7584, 7584, 7584], 148),
                               ([2, 1], [100, 900], 148), ([64, 64, 64, 64, 64, 64], [4464, 4464, 9456, 9456, 14448, 14448], 148)):
        cta, n_ranges = run(tiles, cost, nctas)
        assert cta.min() >= 0 and cta.max() < nctas
        assert (np.diff(cta) >= 0).all()  # contiguous: the walk never goes back to an earlier CTA
        per_tile = np.repeat(np.asarray(cost, np.int64), tiles)
        load = np.bincount(cta, weights=per_tile, minlength=nctas)
        used = min(nctas, int(np.sum(tiles)))
        mean = per_tile.sum() / used
        assert load.max() <= mean + per_tile.max() + 1, (load.max(), mean)
        busy = load[load > 0]
        if int(np.sum(tiles)) >= 4 * nctas:  # enough tiles per CTA for a meaningful lower bound
            assert busy.min() >= mean - per_tile.max() - 1 and len(busy) == nctas
        assert n_ranges >= len(busy) and n_ranges <= len(busy) + len(tiles)


# ---- ONNX initializers without the onnx package (larynx_b200/onnx_weights.py) -------------------------------------
def _pb_varint(x: int) -> bytes:
    x &= (1 << 64) - 1  # negative int64 values travel as their two's complement (10 bytes)
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _pb_field(no: int, payload: bytes) -> bytes:  # length-delimited
    return _pb_varint((no << 3) | 2) + _pb_varint(len(payload)) + payload


def _onnx_tensor(name: str, a: np.ndarray, style: str) -> bytes:
    """TensorProto in one of the encodings exporters use: raw_data, packed float_data, one float_data entry per element."""
    t = b"".join(_pb_varint((1 << 3) | 0) + _pb_varint(int(d)) for d in a.shape) if style != "packed_dims" else \
        _pb_field(1, b"".join(_pb_varint(int(d)) for d in a.shape))
    t += _pb_varint((2 << 3) | 0) + _pb_varint(1)  # FLOAT
    t += _pb_field(8, name.encode())
    flat = np.ascontiguousarray(a, dtype="<f4").reshape(-1)
    if style == "unpacked_floats":
        t += b"".join(_pb_varint((4 << 3) | 5) + flat[i : i + 1].tobytes() for i in range(flat.size))
    elif style == "packed_floats":
        t += _pb_field(4, flat.tobytes())
    else:
        t += _pb_field(9, flat.tobytes())
    return t


def _onnx_model(tensors, extra_int64=None) -> bytes:
    graph = _pb_field(2, b"torch_jit")
    node = _pb_field(4, b"Conv") + _pb_field(1, b"x") + _pb_field(2, b"y")
    graph = _pb_field(1, node) + graph
    styles = ["raw", "packed_floats", "packed_dims", "unpacked_floats"]
    for i, (k, a) in enumerate(tensors.items()):
        graph += _pb_field(5, _onnx_tensor(k, a, styles[i % 4] if a.size <= 64 else styles[i % 3]))
    if extra_int64 is not None:  # shape constants live among the initializers too
        t = _pb_varint((1 << 3) | 0) + _pb_varint(len(extra_int64)) + _pb_varint((2 << 3) | 0) + _pb_varint(7) + _pb_field(8, b"onnx::Reshape_7")
        t += _pb_field(7, b"".join(_pb_varint(int(v)) for v in extra_int64))
        graph += _pb_field(5, t)
    return _pb_varint((1 << 3) | 0) + _pb_varint(7) + _pb_field(2, b"pytorch") + _pb_field(7, graph) + _pb_field(8, _pb_varint((2 << 3) | 0) + _pb_varint(11))


def test_onnx_initializers_are_read_by_name_and_pack_like_the_checkpoint(tmp_path):
    """A vocoder directory that holds only generator.onnx (what released voices ship): the initializers — named like the
    parameters after remove_weight_norm(), hifi_gan/models.py:204-211 — pack to the same blob as the generator.pth of the same
    weights; an export that lost a parameter name fails with that name."""
    from larynx_b200 import onnx_weights, pack, synth

    cfg = synth.hifigan_config("vctk_small")
    sd = synth.hifigan_state_dict(cfg, 7)
    folded = {}
    for k in sd:
        if k.endswith(".weight_g"):
            folded[k[: -len("_g")]] = pack.fold_weight_norm(sd, k[: -len(".weight_g")])
        elif not k.endswith(".weight_v"):
            folded[k] = pack._np(sd[k])
    d = tmp_path / "voc"
    d.mkdir()
    (d / "generator.onnx").write_bytes(_onnx_model({"module." + k if i % 5 == 0 else k: v for i, (k, v) in enumerate(folded.items())},
                                                   extra_int64=[1, -1, 80]))
    ini = onnx_weights.read_initializers(d / "generator.onnx")
    assert ini["onnx::Reshape_7"].dtype == np.int64 and ini["onnx::Reshape_7"].tolist() == [1, -1, 80]
    got = onnx_weights.state_dict(d / "generator.onnx")
    assert set(got) == set(folded)  # float tensors only, `module.` stripped
    for k, v in folded.items():
        assert got[k].dtype == np.float32 and got[k].shape == v.shape and np.array_equal(got[k], v), k
    blob = pack.pack_from_directory(d, "generator", cfg, pack.pack_hifigan)
    assert np.array_equal(blob, pack.pack_hifigan(sd, cfg))

    broken = {k: v for k, v in folded.items() if k != "conv_post.bias"}
    (d / "generator.onnx").write_bytes(_onnx_model(broken))
    with pytest.raises(KeyError, match="conv_post.bias"):
        pack.pack_from_directory(d, "generator", cfg, pack.pack_hifigan)
    with pytest.raises(FileNotFoundError):
        pack.pack_from_directory(tmp_path / "nothing", "generator", cfg, pack.pack_hifigan)
    (d / "generator.onnx").write_bytes(b"\x08\x07")  # a model without a graph
    with pytest.raises(ValueError):
        onnx_weights.read_initializers(d / "generator.onnx")


def _fold_all(sd):
    from larynx_b200 import pack

    out = {}
    for k in sd:
        if k.endswith(".weight_g"):
            out[k[: -len("_g")]] = pack.fold_weight_norm(sd, k[: -len(".weight_g")])
        elif not k.endswith(".weight_v"):
            out[k] = pack._np(sd[k])
    return out


def test_torch_exports_of_the_reference_modules_are_read_back():
    """tests/golden/{glow_tiny,hifigan_tiny}.onnx are the reference's own modules traced by torch's ONNX exporter
    (oracle/make_onnx_fixture.py).  HiFi-GAN keeps every parameter name.  GlowTTS loses the LayerNorm scales/shifts, ActNorm
    logs, the weight-normed `start` convs and the inverted 1x1 matrices to anonymous constants: onnx_weights.glow_state_dict
    recovers them from the execution order — every tensor must equal the checkpoint the export was made from."""
    from larynx_b200 import onnx_weights, synth
    from oracle import make_onnx_fixture as fx

    golden = os.path.join(os.path.dirname(__file__), "golden")
    hcfg = fx.hifigan_tiny_config()
    want = _fold_all(synth.hifigan_state_dict(hcfg, fx.HIFI_SEED))
    got = onnx_weights.state_dict(os.path.join(golden, "hifigan_tiny.onnx"))
    assert set(got) == set(want)
    for k, v in want.items():
        assert got[k].shape == v.shape and np.abs(got[k] - v).max() <= 1e-6 * max(1.0, float(np.abs(v).max())), k

    gcfg = fx.glow_tiny_config()
    truth = synth.glow_state_dict(gcfg, fx.GLOW_SEED)
    got = onnx_weights.glow_state_dict(os.path.join(golden, "glow_tiny.onnx"), gcfg)
    want = _fold_all(truth)
    n_inv = 0
    for k, v in want.items():
        if k.endswith(".weight") and v.shape == (4, 4):  # InvConvNear: the export carries store_inverse()'s matrix
            assert k not in got
            g, v = got[k + "_inv"], np.linalg.inv(v.astype(np.float64)).astype(np.float32)
            n_inv += 1
        else:
            g = got[k]
        assert g.shape == v.shape or g.size == v.size, k
        assert np.abs(g.reshape(-1) - v.reshape(-1)).max() <= 2e-6 * max(1.0, float(np.abs(v).max())), k
    assert n_inv == gcfg["model"]["n_blocks_dec"]
    assert set(got) == {k + "_inv" if (k.endswith(".weight") and want[k].shape == (4, 4)) else k for k in want}
    # a graph that is not a GlowTTS export is refused, not guessed at
    with pytest.raises(ValueError):
        onnx_weights.glow_state_dict(os.path.join(golden, "hifigan_tiny.onnx"), gcfg)
