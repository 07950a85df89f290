#!/usr/bin/env python
"""Regenerate basic_pitch_b200/saved_models/icassp_2022/nmp.bpw from an ONNX export of the network.

Usage: python tools/extract_weights.py [/path/to/nmp.onnx]
Default source: /root/reference/basic_pitch/saved_models/icassp_2022/nmp.onnx (only present in the
build container).  The blob holds tensors only (≈143 KB); see basic_pitch_b200/weights.py for the layout.
"""
import hashlib
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from basic_pitch_b200 import ICASSP_2022_MODEL_PATH, weights  # noqa: E402


def main() -> None:
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/basic_pitch/saved_models/icassp_2022/nmp.onnx"
    w = weights.extract_from_onnx(src)
    blob = weights.pack(w)
    ICASSP_2022_MODEL_PATH.parent.mkdir(parents=True, exist_ok=True)
    ICASSP_2022_MODEL_PATH.write_bytes(blob)
    print(f"wrote {ICASSP_2022_MODEL_PATH} ({len(blob)} B, sha256 {hashlib.sha256(blob).hexdigest()[:16]})")
    for k, v in w.items():
        print(f"  {k:12s} {str(v.shape):16s} sum={float(v.astype('f8').sum()):+.6g} absmax={float(abs(v).max()):.4g}")


if __name__ == "__main__":
    main()
