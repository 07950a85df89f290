"""Debug probe: the three host entry points on the same batch; reports which one deviates and in which frame ranges."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from basic_pitch_b200 import ICASSP_2022_MODEL_PATH, synth
from basic_pitch_b200.inference import Model

model = Model(ICASSP_2022_MODEL_PATH)
lib = model._lib
rng = np.random.default_rng(5)
clips = [synth.tones_clip(float(rng.uniform(0.3, 7.0)), seed=40 + i) for i in range(23)] + [np.zeros(0, np.float32)]
clips += [synth.tones_clip(31.0, seed=99)]
n = len(clips)
flat, offs = model._pack_audio(clips)
frames = [int(lib.bp_num_frames(len(c))) for c in clips]
total = sum(frames)
p = model._params(0.5, 0.3, 11, 11, True, True, True, 0, 88)
ref = model.run_inference_arrays(clips)
refc = np.concatenate([r["contour"] for r in ref])
refn = np.concatenate([r["note"] for r in ref])
foff_ref = np.cumsum([0] + frames)


def run(files_api):
    note, onset = np.zeros((total, 88), np.float32), np.zeros((total, 88), np.float32)
    contour = np.zeros((total, 264), np.float32)
    foff = np.zeros(n + 1, np.int64)
    nt, arrs = model._alloc_notes(n, 4 * total, 64 * total)
    if files_api:
        ptrs = (C.c_void_p * n)(*[c.ctypes.data for c in clips])
        lens = np.array([len(c) for c in clips], np.int64)
        lib.bp_transcribe_files_host(model.handle, ptrs, lens.ctypes.data, n, C.byref(p), note.ctypes.data, onset.ctypes.data,
                                     contour.ctypes.data, foff.ctypes.data, C.byref(nt))
    else:
        lib.bp_transcribe_host(model.handle, flat.ctypes.data, offs.ctypes.data, n, C.byref(p), note.ctypes.data,
                               onset.ctypes.data, contour.ctypes.data, foff.ctypes.data, C.byref(nt))
    return note, contour


for rep in range(2):
    for name, api in (("packed", False), ("files", True)):
        note, contour = run(api)
        bad = np.flatnonzero((contour != refc).any(axis=1))
        badn = np.flatnonzero((note != refn).any(axis=1))
        msg = f"rep {rep} {name}: contour rows differing from run_inference {len(bad)} of {total}; note rows {len(badn)}"
        if len(bad):
            files = sorted({int(np.searchsorted(foff_ref, b, side='right') - 1) for b in bad})
            cols = np.flatnonzero((contour != refc).any(axis=0))
            msg += f"; first {bad[:5]} last {bad[-5:]}; files {files[:12]}; bins {cols[:6]}..{cols[-6:]} ({len(cols)})"
        print(msg, flush=True)
