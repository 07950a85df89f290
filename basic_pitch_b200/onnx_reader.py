"""Schema-less protobuf reader for ONNX model files (no `onnx` / `onnxruntime` needed).

Only what is required to pull the initialisers (weights) and the node list out of the
reference's deployed graph `basic_pitch/saved_models/icassp_2022/nmp.onnx`
(reference: basic_pitch/inference.py:129-137 loads this file with onnxruntime; here we only
read its tensors).  Field numbers follow the public ONNX protobuf schema (onnx.proto3):

  ModelProto  {graph=7}
  GraphProto  {node=1, name=2, initializer=5, input=11, output=12}
  NodeProto   {input=1, output=2, name=3, op_type=4, attribute=5}
  AttributeProto {name=1, f=2, i=3, s=4, t=5, floats=7, ints=8}
  TensorProto {dims=1, data_type=2, float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9}
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np

_WT_VARINT, _WT_I64, _WT_LEN, _WT_I32 = 0, 1, 2, 5


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for every field of one message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == _WT_VARINT:
            v, pos = _varint(buf, pos)
        elif wt == _WT_I64:
            v = buf[pos : pos + 8]
            pos += 8
        elif wt == _WT_LEN:
            ln, pos = _varint(buf, pos)
            v = buf[pos : pos + ln]
            pos += ln
        elif wt == _WT_I32:
            v = buf[pos : pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _packed_varints(v: bytes) -> List[int]:
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _signed64(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


_ONNX_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 11: np.float64, 10: np.float16, 9: np.bool_}


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: List[float] = []
    i32: List[int] = []
    i64: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims.extend(_packed_varints(v) if wt == _WT_LEN else [v])
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            if wt == _WT_LEN:
                floats.extend(struct.unpack(f"<{len(v) // 4}f", v))
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 5:
            i32.extend(_packed_varints(v) if wt == _WT_LEN else [v])
        elif fno == 7:
            i64.extend(_packed_varints(v) if wt == _WT_LEN else [v])
    np_dtype = _ONNX_DTYPES.get(dtype)
    if np_dtype is None:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data_type {dtype}")
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dtype).newbyteorder("<")).astype(np_dtype)
    elif floats:
        arr = np.asarray(floats, dtype=np_dtype)
    elif i64:
        arr = np.asarray([_signed64(x) for x in i64], dtype=np_dtype)
    elif i32:
        arr = np.asarray([_signed64(x) for x in i32], dtype=np_dtype)
    else:
        arr = np.zeros((0,), dtype=np_dtype)
    shape = tuple(int(d) for d in dims)
    return name, arr.reshape(shape) if shape or arr.size == 1 else arr


def _attribute(buf: bytes) -> Tuple[str, object]:
    name = ""
    val: object = None
    ints: List[int] = []
    floats: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _signed64(v)
        elif fno == 4:
            val = bytes(v)
        elif fno == 5:
            val = _tensor(v)[1]
        elif fno == 7:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", v) if wt == _WT_LEN else struct.unpack("<f", v))
        elif fno == 8:
            ints.extend([_signed64(x) for x in _packed_varints(v)] if wt == _WT_LEN else [_signed64(v)])
    if ints:
        val = ints
    elif floats:
        val = floats
    return name, val


class OnnxNode:
    __slots__ = ("op_type", "name", "inputs", "outputs", "attrs")

    def __init__(self) -> None:
        self.op_type = ""
        self.name = ""
        self.inputs: List[str] = []
        self.outputs: List[str] = []
        self.attrs: Dict[str, object] = {}

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        return f"{self.op_type}({', '.join(self.inputs)}) -> {', '.join(self.outputs)} {self.attrs}"


def _node(buf: bytes) -> OnnxNode:
    nd = OnnxNode()
    for fno, _wt, v in _fields(buf):
        if fno == 1:
            nd.inputs.append(bytes(v).decode())
        elif fno == 2:
            nd.outputs.append(bytes(v).decode())
        elif fno == 3:
            nd.name = bytes(v).decode()
        elif fno == 4:
            nd.op_type = bytes(v).decode()
        elif fno == 5:
            k, a = _attribute(v)
            nd.attrs[k] = a
    return nd


def read_onnx(path) -> Tuple[List[OnnxNode], Dict[str, np.ndarray]]:
    """Return (nodes in file order, initialisers by name) of an ONNX file."""
    with open(path, "rb") as fh:
        buf = fh.read()
    graph = None
    for fno, _wt, v in _fields(buf):
        if fno == 7:
            graph = v
    if graph is None:
        raise ValueError(f"{path}: no GraphProto found — not an ONNX model file")
    nodes: List[OnnxNode] = []
    inits: Dict[str, np.ndarray] = {}
    for fno, _wt, v in _fields(graph):
        if fno == 1:
            nodes.append(_node(v))
        elif fno == 5:
            name, arr = _tensor(v)
            inits[name] = arr
    return nodes, inits
