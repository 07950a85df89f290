"""ORACLE (test infrastructure, not product code): CPU restatement of the deployed basic-pitch graph.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product path (`basic_pitch_b200/`) never does.

What it restates: the graph the reference hands to its ML runtime at
reference: basic_pitch/inference.py:156-182 (`Model.predict`), i.e. the contents of
`saved_models/icassp_2022/nmp.onnx` (248 nodes; decoded in SURVEY.md Appendix A), which is the
inference form of
  * `CQT2010v2.call`            reference: basic_pitch/layers/nnaudio.py:623-661
      - `get_cqt_complex`       reference: basic_pitch/layers/nnaudio.py:216-256
      - `downsampling_by_n`     reference: basic_pitch/layers/nnaudio.py:259-284
  * `NormalizedLog.call`        reference: basic_pitch/layers/signal.py:171-185
  * BatchNormalization (folded) reference: basic_pitch/models.py:188-189
  * `HarmonicStacking.call`     reference: basic_pitch/nn.py:69-88 (shifts :51-54)
  * conv stack                  reference: basic_pitch/models.py:241-318

The arithmetic lives in un-vendored third-party runtimes (tensorflow / onnxruntime / tflite /
coremltools, un-pinned ranges in reference pyproject.toml:20-34) that are not installed here, so
this is a restatement, run with torch-CPU convolutions in float32 (default) or float64.

Parity pin: checked against the reference's own golden vector
`tests/resources/vocadito_10/model_output.npz` (reference: tests/test_inference.py:66-70) — see
tests/test_oracle_golden.py and tests/golden/README.md for the measured residuals.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

N_OCTAVES = 9
N_BINS = 309
BINS_PER_OCTAVE = 36
N_FRAMES = 172
N_SAMPLES = 43844
HARMONIC_SHIFTS = (-36, 0, 36, 57, 72, 84, 93, 101)  # round(36*log2(h)), h = 0.5,1,2..7 (nn.py:51-54)
N_CONTOUR_BINS = 264


def _t(a: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def cqt_magnitude(audio: torch.Tensor, w: Dict[str, np.ndarray], dtype=torch.float32) -> torch.Tensor:
    """(B, 43844) -> (B, 172, 309) constant-Q magnitudes (nnaudio.py:623-661)."""
    x = audio.to(dtype)[:, None, :]  # (B,1,L)
    k_re = _t(w["cqt_real"], dtype)[:, None, :]  # (36,1,256)
    k_im = _t(w["cqt_imag"], dtype)[:, None, :]
    lp = _t(w["lowpass"], dtype)[None, None, :]
    octaves = []
    hop = 256
    for o in range(N_OCTAVES):
        p = F.pad(x, (128, 128), mode="reflect")
        re = F.conv1d(p, k_re, stride=hop)  # (B,36,172)
        im = -F.conv1d(p, k_im, stride=hop)
        octaves.insert(0, torch.stack((re, im), dim=-1))  # low octaves first
        if o < N_OCTAVES - 1:
            x = F.conv1d(F.pad(x, (127, 127)), lp, stride=2)
            hop //= 2
    cqt = torch.cat(octaves, dim=1)[:, -N_BINS:]  # (B,309,172,2)
    cqt = cqt * _t(w["cqt_scale"], dtype)[None, :, None, None]
    mag = torch.sqrt((cqt * cqt).sum(-1))  # (B,309,172)
    return mag.transpose(1, 2).contiguous()  # (B,172,309)


def normalized_log(mag: torch.Tensor) -> torch.Tensor:
    """signal.py:171-185 as exported: 10*(ln(p+1e-10)*(1/ln10)), minus per-window min, divided by
    per-window max (0 where the max is 0)."""
    dtype = mag.dtype
    power = mag * mag
    log_power = torch.log(power + torch.tensor(1e-10, dtype=torch.float32).to(dtype))
    log_power = log_power * torch.tensor(0.4342944622039795, dtype=dtype) * 10.0
    mn = log_power.amin(dim=(1, 2), keepdim=True)
    off = log_power - mn
    mx = off.amax(dim=(1, 2), keepdim=True)
    return torch.where(mx == 0, torch.zeros_like(off), off / mx)


def harmonic_stack(y: torch.Tensor) -> torch.Tensor:
    """(B,172,309) -> (B,8,172,264); shifted copies, zero filled (nn.py:69-88)."""
    chans = []
    n = y.shape[-1]
    for s in HARMONIC_SHIFTS:
        if s == 0:
            c = y
        elif s > 0:
            c = F.pad(y[..., s:], (0, s))
        else:
            c = F.pad(y[..., :s], (-s, 0))
        assert c.shape[-1] == n
        chans.append(c[..., :N_CONTOUR_BINS])
    return torch.stack(chans, dim=1)


def forward(audio: np.ndarray, w: Dict[str, np.ndarray], dtype=torch.float32, return_intermediates: bool = False):
    """audio (B,43844) float32 -> dict(note (B,172,88), onset (B,172,88), contour (B,172,264))."""
    a = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    if a.ndim == 3:
        a = a[..., 0]
    assert a.shape[1] == N_SAMPLES, a.shape
    with torch.no_grad():
        mag = cqt_magnitude(a, w, dtype)
        y = normalized_log(mag)
        y = y * _t(w["bn_scale"], dtype) + _t(w["bn_bias"], dtype)
        h = harmonic_stack(y)  # (B,8,172,264)

        def conv(x, name, stride=(1, 1), pad=(0, 0, 0, 0)):
            # pad = (left_f, right_f, top_t, bottom_t)
            return F.conv2d(F.pad(x, pad), _t(w[name + "_w"], dtype), _t(w[name + "_b"], dtype), stride=stride)

        c1 = torch.relu(conv(h, "contour1", pad=(19, 19, 1, 1)))
        contour = torch.sigmoid(conv(c1, "contour2", pad=(2, 2, 2, 2)))  # (B,1,172,264)
        n1 = torch.relu(conv(contour, "note1", stride=(1, 3), pad=(2, 2, 3, 3)))
        note = torch.sigmoid(conv(n1, "note2", pad=(1, 1, 3, 3)))  # (B,1,172,88)
        o1 = torch.relu(conv(h, "onset1", stride=(1, 3), pad=(1, 1, 2, 2)))
        onset = torch.sigmoid(conv(torch.cat((note, o1), dim=1), "onset2", pad=(1, 1, 1, 1)))
    out = {
        "note": note[:, 0].to(torch.float32).numpy(),
        "onset": onset[:, 0].to(torch.float32).numpy(),
        "contour": contour[:, 0].to(torch.float32).numpy(),
    }
    if return_intermediates:
        out["_mag"] = mag.numpy()
        out["_y"] = y.numpy()
        out["_c1"] = c1.numpy()
        out["_n1"] = n1.numpy()
        out["_o1"] = o1.numpy()
    return out


def forward_batched(audio: np.ndarray, w: Dict[str, np.ndarray], dtype=torch.float32, batch: int = 32):
    outs = {"note": [], "onset": [], "contour": []}
    for i in range(0, audio.shape[0], batch):
        o = forward(audio[i : i + batch], w, dtype)
        for k in outs:
            outs[k].append(o[k])
    return {k: np.concatenate(v) for k, v in outs.items()}
