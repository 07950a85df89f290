"""Stand-in runtime: lets the UNMODIFIED reference `basic_pitch.inference.Model` take its ONNX branch
(reference: basic_pitch/inference.py:129-137, 168-182) with the oracle's graph restatement doing the
arithmetic.  Oracle-side only; never importable from the product."""
import numpy as np


def get_available_providers():
    return ["CPUExecutionProvider"]


class InferenceSession:
    _OUT = {"StatefulPartitionedCall:0": "contour", "StatefulPartitionedCall:1": "note", "StatefulPartitionedCall:2": "onset"}

    def __init__(self, path, providers=None):
        import pathlib
        import sys

        root = pathlib.Path(__file__).resolve().parents[3]
        if str(root) not in sys.path:
            sys.path.insert(0, str(root))
        from basic_pitch_b200 import weights

        if not str(path).endswith(".onnx"):
            raise ValueError("not an onnx file")
        self.w = weights.extract_from_onnx(path)
        self.dtype = None

    def run(self, output_names, feeds):
        import torch
        from oracle import model_ref

        (x,) = feeds.values()
        out = model_ref.forward(np.asarray(x, dtype=np.float32), self.w, self.dtype or torch.float32)
        return [out[self._OUT[n]] for n in output_names]
