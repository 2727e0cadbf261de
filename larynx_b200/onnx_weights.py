"""Weights out of an ONNX file without the `onnx` package.

The reference's released voices ship `generator.onnx` (larynx/glow_tts.py:96-100, larynx/hifi_gan.py:101-105 hand it to
onnxruntime); the PyTorch checkpoints the packer normally reads (`generator.pth`) exist only where a voice was trained.
Neither `onnx` nor `onnxruntime` can be installed here, so this module walks the protobuf wire format itself — ModelProto
field 7 (graph) -> GraphProto field 5 (initializer, repeated TensorProto) — and returns the named tensors.  That is enough
for graphs whose initializers kept their parameter names, which is what `torch.onnx.export` does for parameters that
feed Conv / ConvTranspose / Gather / Mul / Add nodes directly (the HiFi-GAN generator after `remove_weight_norm()`).
Parameters that only reach the graph through foldable arithmetic (weight-norm still attached, `exp(-logs)`, the inverted
1x1 flow matrix) are replaced by anonymous constants at export time and cannot be recovered by name: the callers below
then fail with the list of missing names instead of guessing.

NOT validated against a released voice file (none is available offline); validated against files written by
tests/test_pack_and_abi.py with the same wire format.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, Tuple, Union

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


def read_initializers(path) -> Dict[str, np.ndarray]:
    """name -> array (the file's own dtypes) for every initializer of the graph."""
    with open(path, "rb") as f:
        data = memoryview(f.read())
    out: Dict[str, np.ndarray] = {}
    graphs = [v for no, wt, v in _fields(data) if no == 7 and wt == 2]
    if len(graphs) != 1:
        raise ValueError(f"{path}: not an ONNX model (no graph)")
    for no, wt, v in _fields(graphs[0]):
        if no == 5 and wt == 2:
            name, arr = _tensor(v)
            out[name] = arr
    return out


def state_dict(path) -> Dict[str, np.ndarray]:
    """The floating-point initializers as float32, under their names with a leading `module.` removed — what
    pack.load_state_dict returns for a `.pth` checkpoint."""
    sd = {}
    for k, a in read_initializers(path).items():
        if a.dtype.kind == "f":
            sd[k[7:] if k.startswith("module.") else k] = np.ascontiguousarray(a, dtype=np.float32)
    return sd
