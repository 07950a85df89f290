"""Model parameters of the ICASSP-2022 basic-pitch network: extraction, packing, DSP constants.

The reference ships the same network in four containers under
`basic_pitch/saved_models/icassp_2022/` (reference: basic_pitch/__init__.py:74-95) and hands the
file to a third-party runtime (reference: basic_pitch/inference.py:78-154).  This build reads the
tensors out of the ONNX container with `onnx_reader` (no runtime involved) and packs them into a
flat little-endian blob ("BPW1") which is what the C-ABI library consumes (`bp_model_create`,
include/bp_b200.h).  The packed default lives next to this file in
`saved_models/icassp_2022/nmp.bpw`; `tools/extract_weights.py` regenerates it.

Tensor names in the blob (all float32):
  cqt_real, cqt_imag   (36,256)  top-octave CQT kernels           (reference: layers/nnaudio.py:158-213,576-600)
  lowpass              (256,)    half-band FIR for the octave chain (reference: layers/nnaudio.py:45-76,534-538)
  cqt_scale            (309,)    sqrt(kernel length) per bin        (reference: layers/nnaudio.py:590-593,650)
  bn_scale, bn_bias    (1,)      folded BatchNorm after the log     (reference: models.py:188-189)
  contour1_w/b (8,8,3,39)/(8,)   contour2_w/b (1,8,5,5)/(1,)        (reference: models.py:241-262)
  note1_w/b (32,1,7,7)/(32,)     note2_w/b (1,32,7,3)/(1,)          (reference: models.py:270-290)
  onset1_w/b (32,8,5,5)/(32,)    onset2_w/b (1,33,3,3)/(1,)         (reference: models.py:295-313)
"""
from __future__ import annotations

import pathlib
import struct
from typing import Dict, Union

import numpy as np

from .onnx_reader import read_onnx

MAGIC = b"BPW1"

_CONV_SHAPES = {
    (8, 8, 3, 39): "contour1",
    (1, 8, 5, 5): "contour2",
    (32, 1, 7, 7): "note1",
    (1, 32, 7, 3): "note2",
    (32, 8, 5, 5): "onset1",
    (1, 33, 3, 3): "onset2",
}

EXPECTED_SHAPES = {
    "cqt_real": (36, 256),
    "cqt_imag": (36, 256),
    "lowpass": (256,),
    "cqt_scale": (309,),
    "bn_scale": (1,),
    "bn_bias": (1,),
    "contour1_w": (8, 8, 3, 39),
    "contour1_b": (8,),
    "contour2_w": (1, 8, 5, 5),
    "contour2_b": (1,),
    "note1_w": (32, 1, 7, 7),
    "note1_b": (32,),
    "note2_w": (1, 32, 7, 3),
    "note2_b": (1,),
    "onset1_w": (32, 8, 5, 5),
    "onset1_b": (32,),
    "onset2_w": (1, 33, 3, 3),
    "onset2_b": (1,),
}


def extract_from_onnx(path: Union[str, pathlib.Path]) -> Dict[str, np.ndarray]:
    """Pull the network parameters out of an ONNX export of the basic-pitch graph.

    Tensors are identified structurally (by shape and by which node consumes them), not by
    their tf2onnx-generated names, so a re-export of the same architecture also loads.
    """
    nodes, inits = read_onnx(path)
    out: Dict[str, np.ndarray] = {}
    neg_inputs = {n.inputs[0] for n in nodes if n.op_type == "Neg"}
    for n in nodes:
        if n.op_type != "Conv" or n.inputs[1] not in inits:
            continue
        w = inits[n.inputs[1]]
        shp = tuple(w.shape)
        if shp in _CONV_SHAPES:
            key = _CONV_SHAPES[shp]
            out[key + "_w"] = np.ascontiguousarray(w, dtype=np.float32)
            b = inits[n.inputs[2]] if len(n.inputs) > 2 else np.zeros(shp[0], np.float32)
            out[key + "_b"] = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
        elif shp == (36, 1, 1, 256):
            key = "cqt_imag" if n.outputs[0] in neg_inputs else "cqt_real"
            out[key] = np.ascontiguousarray(w.reshape(36, 256), dtype=np.float32)
        elif shp == (1, 1, 1, 256):
            out["lowpass"] = np.ascontiguousarray(w.reshape(256), dtype=np.float32)
    # sqrt(length) vector: the only (309,1,1) initialiser; BN scalars: the Mul/Add pair that
    # follows the log-normalisation (nodes named batch_normalization/...).
    for name, arr in inits.items():
        if tuple(arr.shape) == (309, 1, 1):
            out["cqt_scale"] = np.ascontiguousarray(arr.reshape(309), dtype=np.float32)
    produced = {}
    for n in nodes:
        for o in n.outputs:
            produced[o] = n
    for n in nodes:
        if n.op_type == "Add" and "batch_normalization" in n.outputs[0] and "cq_t" not in n.outputs[0]:
            consts = [inits[i] for i in n.inputs if i in inits and inits[i].size == 1]
            mul = [produced[i] for i in n.inputs if i in produced and produced[i].op_type == "Mul"]
            if consts and mul:
                out["bn_bias"] = consts[0].astype(np.float32).reshape(1)
                mc = [inits[i] for i in mul[0].inputs if i in inits and inits[i].size == 1]
                out["bn_scale"] = mc[0].astype(np.float32).reshape(1)
    validate(out, source=str(path))
    return out


def validate(w: Dict[str, np.ndarray], source: str = "weights") -> None:
    for k, shp in EXPECTED_SHAPES.items():
        if k not in w:
            raise ValueError(f"{source}: tensor {k!r} not found — not a basic-pitch (ICASSP 2022) graph")
        if tuple(w[k].shape) != shp:
            raise ValueError(f"{source}: tensor {k!r} has shape {tuple(w[k].shape)}, expected {shp}")
        if w[k].dtype != np.float32:
            raise ValueError(f"{source}: tensor {k!r} must be float32")
        if not np.all(np.isfinite(w[k])):
            raise ValueError(f"{source}: tensor {k!r} has non-finite values")


def pack(w: Dict[str, np.ndarray]) -> bytes:
    """Serialise to the BPW1 blob: magic, u32 count, then per tensor
    u32 name_len, name bytes (padded to 4), u32 ndim, u32 dims[ndim], f32 data[]."""
    validate(w)
    parts = [MAGIC, struct.pack("<I", len(EXPECTED_SHAPES))]
    for k in EXPECTED_SHAPES:
        name = k.encode()
        pad = (-len(name)) % 4
        a = np.ascontiguousarray(w[k], dtype="<f4")
        parts.append(struct.pack("<I", len(name)))
        parts.append(name + b"\0" * pad)
        parts.append(struct.pack("<I", a.ndim))
        parts.append(struct.pack(f"<{a.ndim}I", *a.shape))
        parts.append(a.tobytes())
    return b"".join(parts)


def unpack(blob: bytes) -> Dict[str, np.ndarray]:
    if blob[:4] != MAGIC:
        raise ValueError("not a BPW1 weight blob")
    (count,) = struct.unpack_from("<I", blob, 4)
    pos = 8
    out: Dict[str, np.ndarray] = {}
    for _ in range(count):
        (nl,) = struct.unpack_from("<I", blob, pos)
        pos += 4
        name = blob[pos : pos + nl].decode()
        pos += nl + ((-nl) % 4)
        (nd,) = struct.unpack_from("<I", blob, pos)
        pos += 4
        dims = struct.unpack_from(f"<{nd}I", blob, pos)
        pos += 4 * nd
        n = int(np.prod(dims)) if nd else 1
        out[name] = np.frombuffer(blob, dtype="<f4", count=n, offset=pos).reshape(dims).astype(np.float32)
        pos += 4 * n
    validate(out, source="BPW1 blob")
    return out


def load(path: Union[str, pathlib.Path]) -> Dict[str, np.ndarray]:
    """Load parameters from a `.bpw` blob or an `.onnx` export (sniffed by content)."""
    path = pathlib.Path(path)
    with open(path, "rb") as fh:
        head = fh.read(4)
    if head == MAGIC:
        return unpack(path.read_bytes())
    return extract_from_onnx(path)


# --------------------------------------------------------------------------------------
# DSP constants regenerated from first principles (used to cross-check the stored ones).
# --------------------------------------------------------------------------------------

SAMPLE_RATE = 22050
BINS_PER_OCTAVE = 36
N_CQT_BINS = 309
FMIN = 27.5
N_FFT = 256


def dsp_constants() -> Dict[str, np.ndarray]:
    """Top-octave CQT kernels, the half-band low-pass and the sqrt(length) scaling, derived
    from the constant-Q definitions the reference layer uses (reference:
    layers/nnaudio.py:530-600 `CQT2010v2.build`, :158-213 `create_cqt_kernels`, :45-76
    `create_lowpass_filter`)."""
    import scipy.signal

    q = 1.0 / (2.0 ** (1.0 / BINS_PER_OCTAVE) - 1.0)
    n_octaves = int(np.ceil(N_CQT_BINS / BINS_PER_OCTAVE))
    remainder = N_CQT_BINS % BINS_PER_OCTAVE
    fmin_top = FMIN * 2 ** (n_octaves - 1)
    top = (remainder - 1) if remainder else (BINS_PER_OCTAVE - 1)
    fmax_top = fmin_top * 2 ** (top / BINS_PER_OCTAVE)
    fmin_top = fmax_top / 2 ** (1 - 1 / BINS_PER_OCTAVE)

    freqs_top = fmin_top * 2.0 ** (np.arange(BINS_PER_OCTAVE) / float(BINS_PER_OCTAVE))
    kern = np.zeros((BINS_PER_OCTAVE, N_FFT), dtype=np.complex64)
    for k, f in enumerate(freqs_top):
        ln = np.ceil(q * SAMPLE_RATE / f)
        start = int(np.ceil(N_FFT / 2.0 - ln / 2.0)) - int(ln % 2)
        n = np.r_[-ln // 2 : ln // 2]
        sig = scipy.signal.get_window("hann", int(ln), fftbins=True) * np.exp(n * 1j * 2 * np.pi * f / SAMPLE_RATE) / ln
        kern[k, start : start + int(ln)] = sig / np.linalg.norm(sig, 1)

    lowpass = scipy.signal.firwin2(256, [0.0, 0.5 / 1.001, 0.5 * 1.001, 1.0], [1.0, 1.0, 0.0, 0.0])
    freqs = FMIN * 2.0 ** (np.arange(N_CQT_BINS) / float(BINS_PER_OCTAVE))
    lengths = np.ceil(q * SAMPLE_RATE / freqs)
    return {
        "cqt_real": kern.real.astype(np.float32),
        "cqt_imag": kern.imag.astype(np.float32),
        "lowpass": lowpass.astype(np.float32),
        "cqt_scale": np.sqrt(lengths.astype(np.float32)),
        "n_octaves": np.int64(n_octaves),
    }
