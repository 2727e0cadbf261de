"""Weights out of an ONNX file without the `onnx` package.

The reference's released voices ship `generator.onnx` (larynx/glow_tts.py:96-100, larynx/hifi_gan.py:101-105 hand it to
onnxruntime); the PyTorch checkpoints the packer normally reads (`generator.pth`) exist only where a voice was trained.
Neither `onnx` nor `onnxruntime` can be installed here, so this module walks the protobuf wire format itself — ModelProto
field 7 (graph) -> GraphProto field 5 (initializer, repeated TensorProto) — and returns the named tensors.  Parameters that
feed Conv / ConvTranspose / Gather nodes directly keep their names in a `torch.onnx.export` (all of the HiFi-GAN generator after
`remove_weight_norm()`: `state_dict`).  GlowTTS loses four kinds of parameters to anonymous constants when its modules are
traced; `glow_state_dict` recovers them from the graph's execution order, which the module code fixes.

Validated against (a) files written by tests/test_pack_and_abi.py in every tensor encoding of the wire format and (b)
tests/golden/{glow_tiny,hifigan_tiny}.onnx, the reference's OWN modules traced by torch 2.11's TorchScript exporter
(oracle/make_onnx_fixture.py; at full size the recovered GlowTTS weights equal the checkpoint to 1.2e-7).  NOT validated against a
released voice file — none is available offline, and those were exported by a 2021 torch whose constant names differ (handled:
`123`, `onnx::Mul_123`, `/scope/Constant_output_0`) and possibly whose folding differs (not known: anything unexpected raises).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Tuple, Union

import numpy as np

_FLOAT, _INT32, _INT64, _FLOAT16, _DOUBLE = 1, 6, 7, 10, 11
_DTYPES = {_FLOAT: np.dtype("<f4"), _INT32: np.dtype("<i4"), _INT64: np.dtype("<i8"), _FLOAT16: np.dtype("<f2"), _DOUBLE: np.dtype("<f8")}


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, Union[int, memoryview]]]:
    """(field number, wire type, value) of one message; length-delimited values are views into `buf` (no copies)."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        no, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield no, wt, v
        elif wt == 1:
            yield no, wt, struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            if pos + n > len(buf):
                raise ValueError("truncated field")
            yield no, wt, buf[pos : pos + n]
            pos += n
        elif wt == 5:
            yield no, wt, struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")


def _signed(x: int) -> int:
    """int32 / int64 values are written as 64-bit two's complement varints."""
    return x - (1 << 64) if x >= (1 << 63) else x


def _packed_varints(v: memoryview):
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    """One TensorProto: dims = 1, data_type = 2, float_data = 4, int32_data = 5, int64_data = 7, name = 8, raw_data = 9,
    double_data = 10, external_data = 13, data_location = 14."""
    dims, dtype, name, raw = [], _FLOAT, "", None
    typed: Dict[int, list] = {4: [], 5: [], 7: [], 10: []}
    for no, wt, v in _fields(buf):
        if no == 1:
            dims.extend(_packed_varints(v) if wt == 2 else [v])
        elif no == 2:
            dtype = int(v)
        elif no == 8:
            name = bytes(v).decode("utf-8")
        elif no == 9:
            raw = v
        elif no in typed:
            if wt == 2:  # packed
                if no == 4:
                    typed[no].append(np.frombuffer(v, dtype="<f4"))
                elif no == 10:
                    typed[no].append(np.frombuffer(v, dtype="<f8"))
                else:
                    typed[no].append(np.array(_packed_varints(v), dtype=np.int64))
            elif no == 4:
                typed[no].append(np.array([v], dtype="<u4").view("<f4"))
            elif no == 10:
                typed[no].append(np.array([v], dtype="<u8").view("<f8"))
            else:
                typed[no].append(np.array([_signed(v)], dtype=np.int64))
        elif no in (13, 14) and (no == 13 or int(v) == 1):
            raise ValueError(f"tensor '{name}' keeps its data in an external file: not supported")
    if dtype not in _DTYPES:
        raise ValueError(f"tensor '{name}': unsupported ONNX data type {dtype}")
    shape = tuple(int(d) for d in dims)
    if raw is not None:
        arr = np.frombuffer(raw, dtype=_DTYPES[dtype])
    elif dtype == _FLOAT:
        arr = np.concatenate(typed[4]) if typed[4] else np.zeros(0, "<f4")
    elif dtype == _DOUBLE:
        arr = np.concatenate(typed[10]) if typed[10] else np.zeros(0, "<f8")
    elif dtype == _INT64:
        arr = np.concatenate(typed[7]) if typed[7] else np.zeros(0, np.int64)
    elif dtype == _INT32:
        arr = (np.concatenate(typed[5]) if typed[5] else np.zeros(0, np.int64)).astype("<i4")
    else:  # FLOAT16 without raw_data: the bit patterns travel in int32_data
        arr = (np.concatenate(typed[5]) if typed[5] else np.zeros(0, np.int64)).astype("<u2").view("<f2")
    n = int(np.prod(shape)) if shape else 1
    if arr.size != n:
        raise ValueError(f"tensor '{name}': {arr.size} elements for shape {shape}")
    return name, arr.reshape(shape)


def _node(buf: memoryview) -> Tuple[str, List[str], List[str], Dict[str, np.ndarray]]:
    """One NodeProto: input = 1, output = 2, op_type = 4, attribute = 5 (AttributeProto: name = 1, t = 5)."""
    op, ins, outs, tensors = "", [], [], {}
    for no, wt, v in _fields(buf):
        if no == 1 and wt == 2:
            ins.append(bytes(v).decode("utf-8"))
        elif no == 2 and wt == 2:
            outs.append(bytes(v).decode("utf-8"))
        elif no == 4 and wt == 2:
            op = bytes(v).decode("utf-8")
        elif no == 5 and wt == 2:
            aname, t = "", None
            for ano, awt, av in _fields(v):
                if ano == 1 and awt == 2:
                    aname = bytes(av).decode("utf-8")
                elif ano == 5 and awt == 2:
                    t = av
            if t is not None:
                tensors[aname] = _tensor(t)[1]
    return op, ins, outs, tensors


def read_graph(path) -> Tuple[Dict[str, np.ndarray], List[Tuple[str, List[str], List[str]]]]:
    """(constants, nodes): every initializer and every `Constant` node's value under its tensor name, and the graph's nodes
    (op type, input names, output names) in file order — which is execution order (ONNX graphs are topologically sorted)."""
    with open(path, "rb") as f:
        data = memoryview(f.read())
    graphs = [v for no, wt, v in _fields(data) if no == 7 and wt == 2]
    if len(graphs) != 1:
        raise ValueError(f"{path}: not an ONNX model (no graph)")
    consts: Dict[str, np.ndarray] = {}
    nodes = []
    for no, wt, v in _fields(graphs[0]):
        if no == 5 and wt == 2:
            name, arr = _tensor(v)
            consts[name] = arr
        elif no == 1 and wt == 2:
            op, ins, outs, tensors = _node(v)
            if op == "Constant" and "value" in tensors and outs:
                consts[outs[0]] = tensors["value"]
            nodes.append((op, ins, outs))
    return consts, nodes


def read_initializers(path) -> Dict[str, np.ndarray]:
    """name -> array (the file's own dtypes) for every initializer of the graph (and every Constant node's value)."""
    return read_graph(path)[0]


def state_dict(path) -> Dict[str, np.ndarray]:
    """The floating-point initializers that kept a parameter name, as float32, with a leading `module.` removed — what
    pack.load_state_dict returns for a `.pth` checkpoint."""
    sd = {}
    for k, a in read_initializers(path).items():
        if a.dtype.kind == "f" and not _anonymous(k):
            sd[k[7:] if k.startswith("module.") else k] = np.ascontiguousarray(a, dtype=np.float32)
    return sd


def _anonymous(name: str) -> bool:
    """Exporters name folded constants `onnx::Mul_123` (torch >= 1.12), just `123` (older) or after the node that makes them
    (`/decoder/flows.34/Constant_7_output_0`); parameters keep their dotted attribute paths."""
    return name.startswith("onnx::") or name.startswith("/") or name.isdigit() or "." not in name


def glow_state_dict(path, config: dict) -> Dict[str, np.ndarray]:
    """GlowTTS weights from `generator.onnx` in the key names pack.pack_glow reads.

    Parameters that feed their node directly keep their names (convs of the encoder and of the WN stacks after
    `remove_weight_norm`, embeddings, biases).  Four kinds are folded into anonymous constants when the reference's modules
    are traced, and are recovered here from the graph's EXECUTION ORDER, which the module code fixes:
      * LayerNorm gamma / beta (`x * gamma.view(1,-1,1) + beta.view(1,-1,1)`, glow_tts/layers.py:19-28): the Mul / Add nodes
        with an anonymous [1, C, 1] constant, in the order prenet 0..2, then per encoder layer norm_layers_1, norm_layers_2
        (attentions.py:62-74), then the duration predictor's two (models.py:39-49);
      * ActNorm `logs` (`exp(-logs)`, layers.py:182-199): the Exp nodes fed by an anonymous [1, 2*mel, 1] constant holding
        -logs; the decoder runs `reversed(flows)` (models.py:191-209), so the k-th one belongs to block n_blocks-1-k;
      * `CouplingBlock.start`, which keeps its weight norm (attentions.py:119-142): Conv nodes whose weight is an anonymous
        [hidden, mel, 1] constant, same block order;
      * the inverted 1x1 matrices of `store_inverse()` (layers.py:238-275): Conv nodes with an [n_split, n_split, 1, 1] weight,
        returned as `decoder.flows.N.weight_inv`.
    Counts and shapes are checked against `config`; anything else raises."""
    m = config["model"]
    H = m.get("hidden_channels_enc") or m["hidden_channels"]
    Hd = m.get("hidden_channels_dec") or m["hidden_channels"]
    Fdp, nb, ns = m["filter_channels_dp"], m["n_blocks_dec"], m["n_split"]
    half = config["audio"]["mel_channels"] * m["n_sqz"] // 2
    consts, nodes = read_graph(path)
    emb = [k for k in consts if k.endswith("encoder.emb.weight")]
    if len(emb) != 1:
        raise ValueError(f"{path}: no initializer named *encoder.emb.weight — not a GlowTTS export with parameter names")
    prefix = emb[0][: -len("encoder.emb.weight")]
    sd: Dict[str, np.ndarray] = {}
    for k, a in consts.items():
        if a.dtype.kind == "f" and k.startswith(prefix) and not _anonymous(k[len(prefix):]):
            sd[k[len(prefix):]] = np.ascontiguousarray(a, dtype=np.float32)

    def anon(name):
        a = consts.get(name)
        return a if a is not None and a.dtype.kind == "f" and _anonymous(name[len(prefix):] if name.startswith(prefix) else name) else None

    gammas, betas, neglogs, starts, invs = [], [], [], [], []
    for op, ins, _ in nodes:
        if op in ("Mul", "Add"):
            c = [a for a in (anon(i) for i in ins) if a is not None and a.ndim == 3 and a.shape[0] == 1 and a.shape[2] == 1 and a.shape[1] in (H, Fdp)]
            if len(c) == 1:
                (gammas if op == "Mul" else betas).append(c[0].reshape(-1))
        elif op == "Exp" and len(ins) == 1:
            a = anon(ins[0])
            if a is not None and a.shape == (1, 2 * half, 1):
                neglogs.append(a)
        elif op == "Conv" and len(ins) >= 2:
            a = anon(ins[1])
            if a is not None and a.shape == (Hd, half, 1):
                starts.append(a)
            elif a is not None and a.shape == (ns, ns, 1, 1):
                invs.append(a.reshape(ns, ns))
    ln_names = [f"encoder.pre.norm_layers.{i}" for i in range(3)]
    for i in range(m["n_layers_enc"]):
        ln_names += [f"encoder.encoder.norm_layers_1.{i}", f"encoder.encoder.norm_layers_2.{i}"]
    ln_names += ["encoder.proj_w.norm_1", "encoder.proj_w.norm_2"]
    want_c = [H] * (len(ln_names) - 2) + [Fdp, Fdp]
    need_ln = [n for n in ln_names if n + ".gamma" not in sd]
    if need_ln:
        if len(need_ln) != len(ln_names) or [len(g) for g in gammas] != want_c or [len(b) for b in betas] != want_c:
            raise ValueError(f"{path}: cannot recover the LayerNorm parameters: {len(gammas)} scale / {len(betas)} shift constants in the "
                             f"graph for {len(ln_names)} LayerNorms")
        for n, g, b in zip(ln_names, gammas, betas):
            sd[n + ".gamma"], sd[n + ".beta"] = np.ascontiguousarray(g, np.float32), np.ascontiguousarray(b, np.float32)

    def per_block(found, what, key, conv):
        missing = [b for b in range(nb) if key(b) not in sd and key(b).replace(".weight", ".weight_v") not in sd]
        if not missing:
            return
        if len(missing) != nb or len(found) != nb:
            raise ValueError(f"{path}: cannot recover {what}: {len(found)} candidates in the graph for {nb} flow blocks")
        for k, a in enumerate(found):  # the decoder runs the blocks last to first
            sd[key(nb - 1 - k)] = np.ascontiguousarray(conv(a), dtype=np.float32)

    per_block(neglogs, "ActNorm.logs", lambda b: f"decoder.flows.{3 * b}.logs", lambda a: -a)
    per_block(starts, "CouplingBlock.start.weight", lambda b: f"decoder.flows.{3 * b + 2}.start.weight", lambda a: a)
    if any(f"decoder.flows.{3 * b + 1}.weight" not in sd for b in range(nb)):
        if len(invs) != nb:
            raise ValueError(f"{path}: cannot recover the inverted 1x1 flow matrices: {len(invs)} candidates for {nb} blocks")
        for k, a in enumerate(invs):
            sd[f"decoder.flows.{3 * (nb - 1 - k) + 1}.weight_inv"] = np.ascontiguousarray(a, dtype=np.float32)
    return sd
