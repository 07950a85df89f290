"""CPU: the C-ABI library loads and exports every symbol include/bp_b200.h declares (no compute calls)."""
import ctypes
import pathlib
import re

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "bp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    from basic_pitch_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bp_b200.h but not exported by libbp_b200.so"
    assert sorted(_lib.EXPORTS) == names


def test_geometry_helpers_match_reference_windowing(golden_dir):
    """bp_num_windows / bp_num_frames vs counts produced by the reference's own windowing code
    (fixture from oracle/make_golden.py; reference: inference.py:194-219, 247-279)."""
    from basic_pitch_b200 import _lib

    lib = _lib.load()
    z = np.load(golden_dir / "host_cases.npz")
    for n, nw, nf in zip(z["lens"], z["n_windows"], z["n_frames"]):
        assert lib.bp_num_windows(int(n)) == int(nw), n
        assert lib.bp_num_frames(int(n)) == int(nf), n


def test_fails_loudly_without_gpu_or_with_bad_blob():
    import torch

    from basic_pitch_b200 import ICASSP_2022_MODEL_PATH, _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    with pytest.raises(_lib.BpError) as e:
        lib.bp_model_create(b"nope" + b"\0" * 60, 64, 0, ctypes.byref(h))
    assert e.value.code == _lib.BP_E_INVALID
    if not torch.cuda.is_available():
        blob = ICASSP_2022_MODEL_PATH.read_bytes()
        with pytest.raises(_lib.BpError) as e:
            lib.bp_model_create(blob, len(blob), 0, ctypes.byref(h))
        assert e.value.code == _lib.BP_E_CUDA
        assert "no CPU path" in str(e.value)


def test_default_decode_params():
    from basic_pitch_b200 import _lib

    lib = _lib.load()
    p = _lib.DecodeParams()
    lib.bp_default_decode_params(ctypes.byref(p))
    assert (p.onset_thresh, p.frame_thresh, p.min_note_len, p.energy_tol) == (0.5, 0.3, 11, 11)
    assert (p.infer_onsets, p.melodia_trick, p.include_pitch_bends, p.min_pitch_idx, p.max_pitch_idx) == (1, 1, 1, 0, 88)
