"""Inference API — drop-in for reference: basic_pitch/inference.py, backed by libbp_b200.so.

Same public names, argument meaning, defaults, return types and error behaviour as the reference
module (`Model`, `predict`, `predict_and_save`, `run_inference`, `window_audio_file`,
`get_audio_input`, `unwrap_output`, `OutputExtensions`, `verify_*`, `build_output_path`,
`save_note_events`, `DEFAULT_*`).  What differs is what sits behind `Model`: instead of dispatching
to TensorFlow / CoreML / TFLite / onnxruntime (reference: inference.py:78-182) there is one runtime —
the hand-written sm_100a kernels in `csrc/` reached through the C ABI of include/bp_b200.h — and no
CPU fallback.  Batch entry points (`Model.transcribe_arrays`, `predict_batch`) are additions.
"""
from __future__ import annotations

import csv
import ctypes as C
import enum
import json
import os
import pathlib
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import ICASSP_2022_MODEL_PATH, _lib, weights
from . import note_creation as infer
from .audio_io import load_audio, load_audio_device
from .constants import (
    ANNOTATIONS_FPS,
    AUDIO_N_SAMPLES,
    AUDIO_SAMPLE_RATE,
    AUDIO_WINDOW_LENGTH,
    FFT_HOP,
    N_FREQ_BINS_CONTOURS,
    N_FREQ_BINS_NOTES,
)

DEFAULT_ONSET_THRESHOLD = 0.5
DEFAULT_FRAME_THRESHOLD = 0.3
DEFAULT_MINIMUM_NOTE_LENGTH_MS = 127.7
DEFAULT_MINIMUM_MIDI_TEMPO = 120
DEFAULT_SONIFICATION_SAMPLERATE = 44100
DEFAULT_OVERLAPPING_FRAMES = 30
DEFAULT_MIDI_VELOCITY_SCALE = 127

_F32 = np.float32


def _ptr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


class _PinnedBlock:
    """One page-locked host allocation (bp_host_alloc).  NumPy arrays carved out of it keep it alive through their
    `.base` chain; when the last of them dies the block goes back to the pool (or is freed)."""

    __slots__ = ("ptr", "nbytes", "pool")

    def __init__(self, ptr: int, nbytes: int, pool: "_PinnedPool"):
        self.ptr, self.nbytes, self.pool = ptr, nbytes, pool

    def __del__(self):
        pool, ptr = self.pool, self.ptr
        self.ptr = 0
        if ptr:
            try:
                pool._release(ptr, self.nbytes)
            except Exception:  # interpreter shutdown
                pass


class _PinnedPool:
    """Pool of page-locked output buffers: device->host copies into them are asynchronous (~50 GB/s instead of the
    ~10 GB/s of pageable memory) and `cudaHostAlloc` itself (~0.2 s per GB) is paid once, not per call."""

    GRANULE = 32 << 20
    KEEP = 3  # free blocks kept per size class

    def __init__(self, lib):
        self._lib = lib
        self._free: Dict[int, List[int]] = {}

    def array(self, shape: Tuple[int, ...], dtype=np.float32) -> np.ndarray:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        cls = max(1, -(-n // self.GRANULE)) * self.GRANULE
        lst = self._free.get(cls)
        ptr = lst.pop() if lst else self._lib.bp_host_alloc(cls)
        if not ptr:
            raise MemoryError(f"cannot allocate {cls} bytes of page-locked host memory")
        carr = (C.c_byte * max(n, 1)).from_address(ptr)
        carr._block = _PinnedBlock(ptr, cls, self)  # lifetime: carr <- ndarray.base <- every view
        return np.frombuffer(carr, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def _release(self, ptr: int, cls: int) -> None:
        lst = self._free.setdefault(cls, [])
        if len(lst) < self.KEEP:
            lst.append(ptr)
        else:
            self._lib.bp_host_free(ptr)

    def __del__(self):
        for lst in self._free.values():
            for ptr in lst:
                try:
                    self._lib.bp_host_free(ptr)
                except Exception:
                    pass
        self._free = {}


def _default_device() -> int:
    for var in ("BP_B200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    return 0


class Model:
    """A loaded network bound to one B200 (reference: inference.py:71-182).

    `model_path` may be the packed blob shipped with this package (`ICASSP_2022_MODEL_PATH`) or an
    `.onnx` export of the same graph (e.g. the reference's `saved_models/icassp_2022/nmp.onnx`).
    Raises ValueError if the file is not a basic-pitch model (like the reference, inference.py:148-154)
    and `_lib.BpError` / ImportError if the CUDA library or a B200 is missing.
    """

    class MODEL_TYPES(enum.Enum):
        B200 = enum.auto()

    def __init__(self, model_path: Union[pathlib.Path, str], device: Optional[int] = None):
        self.model_type = Model.MODEL_TYPES.B200
        self.model_path = pathlib.Path(model_path)
        try:
            w = weights.load(self.model_path)
        except Exception as e:
            raise ValueError(
                f"File {model_path} cannot be loaded as a basic-pitch model (expected the packed .bpw blob or an "
                f"ONNX export of the ICASSP 2022 graph): {e!r}"
            )
        self._lib = _lib.load()
        blob = weights.pack(w)
        self._h = C.c_void_p()
        self.device = _default_device() if device is None else int(device)
        self._lib.bp_model_create(blob, len(blob), self.device, C.byref(self._h))
        self._pinned = _PinnedPool(self._lib)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(self, "_lib", None) is not None:
            try:
                self._lib.bp_model_destroy(h)
            except Exception:
                pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def set_path(self, path: int) -> None:
        """0 = FP32 FFMA kernels everywhere (on-device accuracy reference), 1 = tcgen05 tensor-core kernels with fused
        epilogues (default), 2 = tcgen05 kernels keeping the 8-channel contour activations (for activation-level tests)."""
        self._lib.bp_model_set_path(self._h, int(path))

    @property
    def launch_count(self) -> int:
        return int(self._lib.bp_model_launch_count(self._h))

    # ------------------------------------------------------------------ stage 1+2
    def predict(self, x: np.ndarray) -> Dict[str, np.ndarray]:
        """(B, 43844, 1) or (B, 43844) float32 -> {"note","onset","contour"} (reference: inference.py:156-182)."""
        x = np.asarray(x)
        if x.ndim == 3 and x.shape[2] == 1:
            x = x[:, :, 0]
        if x.ndim != 2 or x.shape[1] != AUDIO_N_SAMPLES:
            raise ValueError(f"expected audio of shape (B, {AUDIO_N_SAMPLES}, 1), got {x.shape}")
        x = np.ascontiguousarray(x, dtype=_F32)
        n = x.shape[0]
        note = np.empty((n, 172, N_FREQ_BINS_NOTES), _F32)
        onset = np.empty((n, 172, N_FREQ_BINS_NOTES), _F32)
        contour = np.empty((n, 172, N_FREQ_BINS_CONTOURS), _F32)
        self._lib.bp_forward_host(self._h, _ptr(x), n, _ptr(note), _ptr(onset), _ptr(contour))
        return {"note": note, "onset": onset, "contour": contour}

    # ------------------------------------------------------------------ whole files
    @staticmethod
    def _pack_audio(audios: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        offs = np.zeros(len(audios) + 1, dtype=np.int64)
        for i, a in enumerate(audios):
            if a.ndim != 1:
                raise ValueError("audio must be mono (1-D)")
            offs[i + 1] = offs[i] + a.shape[0]
        flat = np.empty(max(int(offs[-1]), 1), dtype=_F32)
        for i, a in enumerate(audios):
            flat[offs[i] : offs[i + 1]] = a
        return flat, offs

    def run_inference_arrays(self, audios: Sequence[np.ndarray]) -> List[Dict[str, np.ndarray]]:
        """Windowing + model + unwrap for a batch of mono 22 050 Hz signals (reference: inference.py:282-330)."""
        flat, offs = self._pack_audio(audios)
        n_files = len(audios)
        frames = [int(self._lib.bp_num_frames(int(offs[i + 1] - offs[i]))) for i in range(n_files)]
        total = sum(frames)
        note = np.empty((total, N_FREQ_BINS_NOTES), _F32)
        onset = np.empty((total, N_FREQ_BINS_NOTES), _F32)
        contour = np.empty((total, N_FREQ_BINS_CONTOURS), _F32)
        foff = np.zeros(n_files + 1, dtype=np.int64)
        self._lib.bp_run_inference_host(self._h, _ptr(flat), _ptr(offs), n_files, _ptr(note), _ptr(onset), _ptr(contour), _ptr(foff))
        return [
            {"note": note[foff[i] : foff[i + 1]], "onset": onset[foff[i] : foff[i + 1]], "contour": contour[foff[i] : foff[i + 1]]}
            for i in range(n_files)
        ]

    # ------------------------------------------------------------------ stage 3
    def _params(self, onset_thresh, frame_thresh, min_note_len, energy_tol, infer_onsets, melodia_trick,
                include_pitch_bends, min_pitch_idx, max_pitch_idx) -> _lib.DecodeParams:
        p = _lib.DecodeParams()
        p.onset_thresh, p.frame_thresh = float(onset_thresh), float(frame_thresh)
        p.min_note_len, p.energy_tol = int(min_note_len), int(energy_tol)
        p.infer_onsets, p.melodia_trick = int(bool(infer_onsets)), int(bool(melodia_trick))
        p.include_pitch_bends = int(bool(include_pitch_bends))
        p.min_pitch_idx, p.max_pitch_idx = int(min_pitch_idx), int(max_pitch_idx)
        return p

    @staticmethod
    def _alloc_notes(n_files: int, note_cap: int, bend_cap: int):
        arrs = {
            "note_off": np.zeros(n_files + 1, np.int32),
            "start": np.empty(note_cap, np.int32),
            "end": np.empty(note_cap, np.int32),
            "pitch": np.empty(note_cap, np.int32),
            "amp": np.empty(note_cap, _F32),
            "bend_off": np.zeros(note_cap + 1, np.int32),
            "bends": np.empty(max(bend_cap, 1), np.int32),
        }
        n = _lib.Notes()
        n.note_capacity, n.bend_capacity = note_cap, bend_cap
        n.note_off, n.start_frame, n.end_frame = _ptr(arrs["note_off"]), _ptr(arrs["start"]), _ptr(arrs["end"])
        n.pitch_midi, n.amplitude = _ptr(arrs["pitch"]), _ptr(arrs["amp"])
        n.bend_off, n.bends = _ptr(arrs["bend_off"]), _ptr(arrs["bends"])
        return n, arrs

    @staticmethod
    def _split_notes(arrs, n_files: int) -> List[Dict[str, np.ndarray]]:
        out = []
        noff, boff = arrs["note_off"], arrs["bend_off"]
        for i in range(n_files):
            a, b = int(noff[i]), int(noff[i + 1])
            b0 = int(boff[a])
            out.append({
                "start": arrs["start"][a:b].copy(), "end": arrs["end"][a:b].copy(), "pitch": arrs["pitch"][a:b].copy(),
                "amp": arrs["amp"][a:b].copy(), "bend_off": (boff[a : b + 1] - b0).copy(),
                "bends": arrs["bends"][b0 : int(boff[b])].copy(),
            })  # fmt: skip
        return out

    def _with_capacity(self, n_files: int, total_frames: int, call):
        note_cap = max(4096, 2 * total_frames)
        bend_cap = max(65536, 24 * total_frames)
        for _ in range(3):  # at most: notes too small, then bends too small, then success
            notes, arrs = self._alloc_notes(n_files, note_cap, bend_cap)
            try:
                call(notes)
                return arrs
            except _lib.BpError as e:
                if e.code != _lib.BP_E_CAPACITY:
                    raise
                need_n, need_b = C.c_int64(0), C.c_int64(0)
                self._lib.bp_last_required(C.byref(need_n), C.byref(need_b))
                if need_n.value <= note_cap and need_b.value <= bend_cap:
                    raise
                note_cap, bend_cap = max(note_cap, need_n.value), max(bend_cap, need_b.value)
        raise RuntimeError("decode capacity negotiation failed")

    def decode_arrays(self, notes: Sequence[np.ndarray], onsets: Sequence[np.ndarray],
                      contours: Optional[Sequence[np.ndarray]], onset_thresh=0.5, frame_thresh=0.3, min_note_len=11,
                      energy_tol=11, infer_onsets=True, melodia_trick=True, include_pitch_bends=True,
                      min_pitch_idx=0, max_pitch_idx=88) -> List[Dict[str, np.ndarray]]:
        """Posteriorgrams of a batch of files -> note arrays per file (reference: note_creation.py:52-111)."""
        n_files = len(notes)
        foff = np.zeros(n_files + 1, np.int64)
        for i, a in enumerate(notes):
            if a.shape[1:] != (N_FREQ_BINS_NOTES,) or onsets[i].shape != a.shape:
                raise ValueError("note/onset posteriorgrams must be (T, 88) and of equal shape")
            foff[i + 1] = foff[i] + a.shape[0]
        total = int(foff[-1])
        cat = lambda xs, w: np.ascontiguousarray(np.concatenate([np.asarray(x, _F32).reshape(-1, w) for x in xs]) if xs else np.zeros((0, w), _F32), dtype=_F32)  # noqa: E731
        n_all, o_all = cat(list(notes), N_FREQ_BINS_NOTES), cat(list(onsets), N_FREQ_BINS_NOTES)
        if contours is None:
            if include_pitch_bends:
                raise ValueError("pitch bends need the contour posteriorgram")
            c_all = np.zeros((max(total, 1), N_FREQ_BINS_CONTOURS), _F32)
        else:
            c_all = cat(list(contours), N_FREQ_BINS_CONTOURS)
            if c_all.shape[0] != total:
                raise ValueError("contour posteriorgrams must have as many frames as note posteriorgrams")
        p = self._params(onset_thresh, frame_thresh, min_note_len, energy_tol, infer_onsets, melodia_trick,
                         include_pitch_bends, min_pitch_idx, max_pitch_idx)
        arrs = self._with_capacity(
            n_files, total,
            lambda nt: self._lib.bp_decode_host(self._h, _ptr(n_all), _ptr(o_all), _ptr(c_all), _ptr(foff), n_files, C.byref(p), C.byref(nt)),
        )
        return self._split_notes(arrs, n_files)

    def infer_onsets_array(self, onsets: np.ndarray, frames: np.ndarray) -> np.ndarray:
        """reference: note_creation.py:289-311 `get_infered_onsets` (n_diff = 2) -> float64 (T, 88), on the device."""
        o = np.ascontiguousarray(onsets, dtype=_F32)
        f = np.ascontiguousarray(frames, dtype=_F32)
        if o.ndim != 2 or o.shape[1] != N_FREQ_BINS_NOTES or f.shape != o.shape:
            raise ValueError("onsets / frames must be (T, 88) and of equal shape")
        out = np.empty(o.shape, np.float64)
        self._lib.bp_infer_onsets_host(self._h, _ptr(o), _ptr(f), o.shape[0], _ptr(out))
        return out

    def pitch_bends_arrays(self, contours: np.ndarray, start: np.ndarray, end: np.ndarray, pitch: np.ndarray):
        """reference: note_creation.py:182-219 `get_pitch_bends` for given notes -> (bend_off int32[n+1], bends int32[])."""
        c = np.ascontiguousarray(contours, dtype=_F32)
        if c.ndim != 2 or c.shape[1] != N_FREQ_BINS_CONTOURS:
            raise ValueError("contours must be (T, 264)")
        st = np.ascontiguousarray(start, dtype=np.int32)
        en = np.ascontiguousarray(end, dtype=np.int32)
        pi = np.ascontiguousarray(pitch, dtype=np.int32)
        n = len(st)
        off = np.zeros(n + 1, np.int32)
        cap = int(np.maximum(en.astype(np.int64) - st.astype(np.int64), 0).sum())
        bends = np.empty(max(cap, 1), np.int32)
        self._lib.bp_pitch_bends_host(self._h, _ptr(c), c.shape[0], n, _ptr(st), _ptr(en), _ptr(pi), _ptr(off), _ptr(bends), cap)
        return off, bends[: int(off[n])]

    # ------------------------------------------------------------------ the whole path
    def transcribe_arrays(self, audios: Sequence[np.ndarray], onset_thresh=0.5, frame_thresh=0.3, min_note_len=11,
                          energy_tol=11, infer_onsets=True, melodia_trick=True, include_pitch_bends=True,
                          min_pitch_idx=0, max_pitch_idx=88, return_model_output: bool = True, split_notes: bool = True):
        """Audio of a batch of files -> (model outputs | None, note arrays) per file in ONE library call
        (reference: inference.py:431-506 `predict`, minus file I/O and the MIDI object).  With split_notes=False the
        second result is the concatenated arrays of the call (note_off, start, end, pitch, amp, bend_off, bends)."""
        n_files = len(audios)
        # one pointer per file: the library gathers the (pageable) arrays into pinned staging itself, sub-batch by
        # sub-batch, overlapped with the kernels (bp_transcribe_files_host) — no host-side concatenation
        keep = []
        for a in audios:
            if a.ndim != 1:
                raise ValueError("audio must be mono (1-D)")
            keep.append(a if (a.dtype == _F32 and a.flags.c_contiguous) else np.ascontiguousarray(a, dtype=_F32))
        ptrs = (C.c_void_p * max(n_files, 1))(*[a.ctypes.data for a in keep])
        lens = np.fromiter((a.shape[0] for a in keep), dtype=np.int64, count=n_files)
        frames = [int(self._lib.bp_num_frames(int(n))) for n in lens]
        total = sum(frames)
        note = onset = contour = None
        if return_model_output:  # page-locked: the posteriorgrams stream back while later sub-batches compute
            note = self._pinned.array((total, N_FREQ_BINS_NOTES))
            onset = self._pinned.array((total, N_FREQ_BINS_NOTES))
            contour = self._pinned.array((total, N_FREQ_BINS_CONTOURS))
        foff = np.zeros(n_files + 1, np.int64)
        p = self._params(onset_thresh, frame_thresh, min_note_len, energy_tol, infer_onsets, melodia_trick,
                         include_pitch_bends, min_pitch_idx, max_pitch_idx)
        arrs = self._with_capacity(
            n_files, total,
            lambda nt: self._lib.bp_transcribe_files_host(self._h, ptrs, _ptr(lens), n_files, C.byref(p), _ptr(note),
                                                          _ptr(onset), _ptr(contour), _ptr(foff), C.byref(nt)),
        )
        res = self._split_notes(arrs, n_files) if split_notes else arrs
        outs: List[Optional[Dict[str, np.ndarray]]] = [None] * n_files
        if return_model_output:
            fo = foff.tolist()
            outs = [{"note": note[a:b], "onset": onset[a:b], "contour": contour[a:b]} for a, b in zip(fo[:-1], fo[1:])]
        return outs, res, frames


_DEFAULT_MODELS: Dict[Tuple[str, int], Model] = {}


def default_model(model_path: Union[pathlib.Path, str] = ICASSP_2022_MODEL_PATH, device: Optional[int] = None) -> Model:
    """One shared `Model` per (path, device); what `predict(path)` uses when given a path."""
    dev = _default_device() if device is None else int(device)
    key = (str(model_path), dev)
    if key not in _DEFAULT_MODELS:
        _DEFAULT_MODELS[key] = Model(model_path, dev)
    return _DEFAULT_MODELS[key]


# ---------------------------------------------------------------------------------------------
# Windowing helpers (host-side equivalents kept for API compatibility; the hot path does this
# arithmetic on the device, see bp_run_inference_* in include/bp_b200.h).
# ---------------------------------------------------------------------------------------------
def window_audio_file(audio_original: np.ndarray, hop_size: int) -> Iterable[Tuple[np.ndarray, Dict[str, float]]]:
    """reference: inference.py:194-219"""
    for i in range(0, audio_original.shape[0], hop_size):
        window = audio_original[i : i + AUDIO_N_SAMPLES]
        if len(window) < AUDIO_N_SAMPLES:
            window = np.pad(window, pad_width=[[0, AUDIO_N_SAMPLES - len(window)]])
        t_start = float(i) / AUDIO_SAMPLE_RATE
        yield np.expand_dims(window, axis=-1), {"start": t_start, "end": t_start + (AUDIO_N_SAMPLES / AUDIO_SAMPLE_RATE)}


def get_audio_input(audio_path: Union[pathlib.Path, str], overlap_len: int, hop_size: int):
    """reference: inference.py:222-244"""
    assert overlap_len % 2 == 0, f"overlap_length must be even, got {overlap_len}"
    audio_original, _ = load_audio(audio_path, sr=AUDIO_SAMPLE_RATE, mono=True)
    original_length = audio_original.shape[0]
    audio_original = np.concatenate([np.zeros((int(overlap_len / 2),), dtype=np.float32), audio_original])
    for window, window_time in window_audio_file(audio_original, hop_size):
        yield np.expand_dims(window, axis=0), window_time, original_length


def unwrap_output(output: np.ndarray, audio_original_length: int, n_overlapping_frames: int, hop_size: int):
    """reference: inference.py:247-279"""
    if len(output.shape) != 3:
        return None
    n_olap = int(0.5 * n_overlapping_frames)
    if n_olap > 0:
        output = output[:, n_olap:-n_olap, :]
    flat = output.reshape(output.shape[0] * output.shape[1], output.shape[2])
    n_expected_windows = audio_original_length / hop_size
    n_frames_per_window = (AUDIO_WINDOW_LENGTH * ANNOTATIONS_FPS) - n_overlapping_frames
    return flat[: int(n_expected_windows * n_frames_per_window), :]


def run_inference(audio_path: Union[pathlib.Path, str], model_or_model_path: Union[Model, pathlib.Path, str],
                  debug_file: Optional[pathlib.Path] = None) -> Dict[str, np.ndarray]:
    """reference: inference.py:282-330 — returns the unwrapped note / onset / contour posteriorgrams."""
    model = model_or_model_path if isinstance(model_or_model_path, Model) else default_model(model_or_model_path)
    audio, _ = load_audio(audio_path, sr=AUDIO_SAMPLE_RATE, mono=True)
    out = model.run_inference_arrays([audio])[0]
    if debug_file:
        n_overlap = DEFAULT_OVERLAPPING_FRAMES * FFT_HOP
        with open(debug_file, "w") as f:
            hop = AUDIO_N_SAMPLES - n_overlap
            padded = np.concatenate([np.zeros(n_overlap // 2, _F32), audio])
            last = padded[(max(len(padded) - 1, 0) // hop) * hop :][:AUDIO_N_SAMPLES]  # the reference dumps the LAST window
            json.dump({
                "audio_windowed": np.pad(last, (0, AUDIO_N_SAMPLES - len(last))).reshape(1, AUDIO_N_SAMPLES, 1).tolist(),
                "audio_original_length": int(audio.shape[0]),
                "hop_size_samples": AUDIO_N_SAMPLES - n_overlap,
                "overlap_length_samples": n_overlap,
                "unwrapped_output": {k: v.tolist() for k, v in out.items()},
            }, f)  # fmt: skip
    return out


def run_inference_stream(audio_blocks: Iterable[np.ndarray], model_or_model_path: Union["Model", pathlib.Path, str] = ICASSP_2022_MODEL_PATH,
                         windows_per_step: int = 64) -> Iterable[Dict[str, np.ndarray]]:
    """Bounded-memory inference for long recordings (reference: README.md:196-198 — "we recommend streaming the audio of
    the file, processing windows of audio at a time").

    audio_blocks: an iterable of consecutive mono 22 050 Hz float blocks of any sizes (e.g. a file read piecewise).
    Yields dicts {"note", "onset", "contour"} of consecutive unwrapped frames; concatenated they equal
    `run_inference(whole_file)` bit for bit.  At any time the host holds at most `windows_per_step` windows of audio
    (~1.6 s each) plus two windows of held-back frames, the device one step of windows: memory does not grow with the
    length of the recording.  (The final trim of `unwrap_output`, inference.py:247-279, can reach two windows back, so
    the frames of the last two windows are yielded when the stream ends.)"""
    model = model_or_model_path if isinstance(model_or_model_path, Model) else default_model(model_or_model_path)
    n_overlap = DEFAULT_OVERLAPPING_FRAMES * FFT_HOP
    hop = AUDIO_N_SAMPLES - n_overlap
    n_olap = DEFAULT_OVERLAPPING_FRAMES // 2
    per_window = AUDIO_WINDOW_LENGTH * ANNOTATIONS_FPS - DEFAULT_OVERLAPPING_FRAMES  # 142 kept frames per window
    buf = np.zeros(n_overlap // 2, _F32)  # the reference prepends overlap_len / 2 zeros (inference.py:241)
    total = 0  # samples of the recording seen so far
    emitted = 0  # frames yielded so far
    held: List[Dict[str, np.ndarray]] = []  # frames of the newest windows, not yet safe to yield

    def run(windows: np.ndarray) -> None:
        out = model.predict(windows)
        for k in range(windows.shape[0]):
            held.append({key: out[key][k, n_olap : 172 - n_olap] for key in ("note", "onset", "contour")})

    def drain(keep: int):
        nonlocal emitted
        while len(held) > keep:
            w = held.pop(0)
            emitted += per_window
            yield w

    for block in audio_blocks:
        block = np.asarray(block, dtype=_F32).reshape(-1)
        total += block.shape[0]
        buf = np.concatenate([buf, block])
        n_ready = (buf.shape[0] - AUDIO_N_SAMPLES) // hop + 1 if buf.shape[0] >= AUDIO_N_SAMPLES else 0
        while n_ready > 0:
            n = min(n_ready, windows_per_step)
            idx = np.arange(n)[:, None] * hop + np.arange(AUDIO_N_SAMPLES)[None, :]
            run(buf[idx])
            buf = buf[n * hop :]
            n_ready -= n
            yield from drain(2)
    # end of stream: the windows that start inside the remaining samples, zero-padded (inference.py:194-219)
    tail = []
    for i in range(0, buf.shape[0], hop):
        w = buf[i : i + AUDIO_N_SAMPLES]
        tail.append(np.pad(w, (0, AUDIO_N_SAMPLES - w.shape[0])))
    for c0 in range(0, len(tail), windows_per_step):
        run(np.stack(tail[c0 : c0 + windows_per_step]))
    # the trim of unwrap_output: keep int(n_samples / hop * frames_per_window) frames in all
    n_final = int(total / hop * per_window)
    for w in held:
        take = max(0, min(per_window, n_final - emitted))
        emitted += take
        if take > 0:
            yield {k: v[:take] for k, v in w.items()}
    held.clear()


def predict_stream(audio_blocks: Iterable[np.ndarray], model_or_model_path: Union["Model", pathlib.Path, str] = ICASSP_2022_MODEL_PATH,
                   onset_threshold: float = DEFAULT_ONSET_THRESHOLD, frame_threshold: float = DEFAULT_FRAME_THRESHOLD,
                   minimum_note_length: float = DEFAULT_MINIMUM_NOTE_LENGTH_MS, minimum_frequency: Optional[float] = None,
                   maximum_frequency: Optional[float] = None, multiple_pitch_bends: bool = False, melodia_trick: bool = True,
                   midi_tempo: float = DEFAULT_MINIMUM_MIDI_TEMPO, windows_per_step: int = 64):
    """`predict` for a recording delivered block by block: the model runs with bounded memory (`run_inference_stream`), the
    posteriorgrams (1.76 KB per frame, 86 frames per second) are collected on the host and decoded once — the note
    decode is global over a file (reference: note_creation.py:409-509).  Returns (model_output, midi_data, note_events)."""
    model = model_or_model_path if isinstance(model_or_model_path, Model) else default_model(model_or_model_path)
    parts = list(run_inference_stream(audio_blocks, model, windows_per_step))
    keys = ("note", "onset", "contour")
    widths = {"note": N_FREQ_BINS_NOTES, "onset": N_FREQ_BINS_NOTES, "contour": N_FREQ_BINS_CONTOURS}
    model_output = {k: (np.concatenate([p[k] for p in parts]) if parts else np.zeros((0, widths[k]), _F32)) for k in keys}
    min_note_len = int(np.round(minimum_note_length / 1000 * (AUDIO_SAMPLE_RATE / FFT_HOP)))
    midi_data, note_events = infer.model_output_to_notes(
        model_output, onset_thresh=onset_threshold, frame_thresh=frame_threshold, min_note_len=min_note_len,
        min_freq=minimum_frequency, max_freq=maximum_frequency, multiple_pitch_bends=multiple_pitch_bends,
        melodia_trick=melodia_trick, midi_tempo=midi_tempo, model=model,
    )
    return model_output, midi_data, note_events


class OutputExtensions(enum.Enum):
    MIDI = "mid"
    MODEL_OUTPUT_NPZ = "npz"
    MIDI_SONIFICATION = "wav"
    NOTE_EVENTS = "csv"


def verify_input_path(audio_path: Union[pathlib.Path, str]) -> None:
    if not os.path.isfile(audio_path):
        raise ValueError(f"🚨 {audio_path} is not a file path.")
    if not os.path.exists(audio_path):
        raise ValueError(f"🚨 {audio_path} does not exist.")


def verify_output_dir(output_dir: Union[pathlib.Path, str]) -> None:
    if not os.path.isdir(output_dir):
        raise ValueError(f"🚨 {output_dir} is not a directory.")
    if not os.path.exists(output_dir):
        raise ValueError(f"🚨 {output_dir} does not exist.")


def build_output_path(audio_path: Union[pathlib.Path, str], output_directory: Union[pathlib.Path, str],
                      output_type: OutputExtensions) -> pathlib.Path:
    """reference: inference.py:372-406 — `<stem>_basic_pitch.<ext>`, never overwrites."""
    basename, _ = os.path.splitext(os.path.basename(str(audio_path)))
    output_path = pathlib.Path(output_directory) / f"{basename}_basic_pitch.{output_type.value}"
    print(f"\n\n  Creating {output_type.name.lower().replace('_', ' ')}...")
    if output_path.exists():
        raise IOError(f"  🚨 {str(output_path)} already exists and would be overwritten. Skipping output files for {audio_path}.")
    return output_path


def save_note_events(note_events: List[infer.NoteEvent], save_path: Union[pathlib.Path, str]) -> None:
    """reference: inference.py:409-428"""
    with open(save_path, "w") as fhandle:
        writer = csv.writer(fhandle, delimiter=",")
        writer.writerow(["start_time_s", "end_time_s", "pitch_midi", "velocity", "pitch_bend"])
        for start_time, end_time, note_number, amplitude, pitch_bend in note_events:
            row = [start_time, end_time, note_number, int(np.round(DEFAULT_MIDI_VELOCITY_SCALE * amplitude))]
            if pitch_bend:
                row.extend(pitch_bend)
            writer.writerow(row)


def _events_and_midi(model_output, res, n_frames, min_freq, max_freq, multiple_pitch_bends, midi_tempo):
    events = infer.note_events_from_arrays(res, n_frames, include_pitch_bends=True)
    if min_freq is not None or max_freq is not None:  # the reference zeroes these columns of the returned arrays
        lo, hi = infer.frequency_to_column_range(min_freq, max_freq)
        for k in ("note", "onset"):
            model_output[k][:, :lo] = 0
            model_output[k][:, hi:] = 0
    return infer.note_events_to_midi(events, multiple_pitch_bends, midi_tempo), events


def predict(
    audio_path: Union[pathlib.Path, str],
    model_or_model_path: Union[Model, pathlib.Path, str] = ICASSP_2022_MODEL_PATH,
    onset_threshold: float = DEFAULT_ONSET_THRESHOLD,
    frame_threshold: float = DEFAULT_FRAME_THRESHOLD,
    minimum_note_length: float = DEFAULT_MINIMUM_NOTE_LENGTH_MS,
    minimum_frequency: Optional[float] = None,
    maximum_frequency: Optional[float] = None,
    multiple_pitch_bends: bool = False,
    melodia_trick: bool = True,
    debug_file: Optional[pathlib.Path] = None,
    midi_tempo: float = DEFAULT_MINIMUM_MIDI_TEMPO,
):
    """reference: inference.py:431-506 -> (model_output, midi_data, note_events)."""
    print(f"Predicting MIDI for {audio_path}...")
    model = model_or_model_path if isinstance(model_or_model_path, Model) else default_model(model_or_model_path)
    audio, _ = load_audio_device(audio_path, model)  # decode on the host, convert / down-mix / resample on the GPU
    min_note_len = int(np.round(minimum_note_length / 1000 * (AUDIO_SAMPLE_RATE / FFT_HOP)))
    lo, hi = infer.frequency_to_column_range(minimum_frequency, maximum_frequency)
    outs, res, frames = model.transcribe_arrays(
        [audio], onset_thresh=onset_threshold, frame_thresh=frame_threshold, min_note_len=min_note_len,
        melodia_trick=melodia_trick, min_pitch_idx=lo, max_pitch_idx=hi,
    )
    model_output = outs[0]
    midi_data, note_events = _events_and_midi(model_output, res[0], frames[0], minimum_frequency, maximum_frequency,
                                              multiple_pitch_bends, midi_tempo)
    if debug_file:
        with open(debug_file, "w") as f:
            json.dump({
                "audio_original_length": int(audio.shape[0]),
                "unwrapped_output": {k: v.tolist() for k, v in model_output.items()},
                "min_note_length": min_note_len,
                "onset_thresh": onset_threshold,
                "frame_thresh": frame_threshold,
                "estimated_notes": [
                    (float(s), float(e), int(p), float(a), [int(b) for b in pb] if pb else None)
                    for s, e, p, a, pb in note_events
                ],
            }, f)  # fmt: skip
    return model_output, midi_data, note_events


def predict_batch(
    audio: Sequence[Union[np.ndarray, pathlib.Path, str]],
    model_or_model_path: Union[Model, pathlib.Path, str] = ICASSP_2022_MODEL_PATH,
    onset_threshold: float = DEFAULT_ONSET_THRESHOLD,
    frame_threshold: float = DEFAULT_FRAME_THRESHOLD,
    minimum_note_length: float = DEFAULT_MINIMUM_NOTE_LENGTH_MS,
    minimum_frequency: Optional[float] = None,
    maximum_frequency: Optional[float] = None,
    multiple_pitch_bends: bool = False,
    melodia_trick: bool = True,
    midi_tempo: float = DEFAULT_MINIMUM_MIDI_TEMPO,
    return_model_output: bool = True,
    build_midi: bool = True,
    lazy: bool = True,
):
    """`predict` for many clips in one device pass (addition; no reference counterpart).

    Items are paths or mono 22 050 Hz float arrays.  Returns a list of
    (model_output | None, midi_data | None, note_events) in input order.  The model outputs are views of three
    page-locked arrays shared by the batch.  With lazy=True (default) the note events are `NoteEventList`s and the MIDI
    objects `LazyPrettyMIDI`s: both turn into the reference's Python objects when first read; lazy=False builds
    everything before returning."""
    model = model_or_model_path if isinstance(model_or_model_path, Model) else default_model(model_or_model_path)
    audios = [a if isinstance(a, np.ndarray) else load_audio_device(a, model)[0] for a in audio]
    min_note_len = int(np.round(minimum_note_length / 1000 * (AUDIO_SAMPLE_RATE / FFT_HOP)))
    lo, hi = infer.frequency_to_column_range(minimum_frequency, maximum_frequency)
    outs, arrs, frames = model.transcribe_arrays(
        audios, onset_thresh=onset_threshold, frame_thresh=frame_threshold, min_note_len=min_note_len,
        melodia_trick=melodia_trick, min_pitch_idx=lo, max_pitch_idx=hi, return_model_output=return_model_output,
        split_notes=False,
    )
    n = len(audios)
    events = infer.note_events_batch(arrs, n, include_pitch_bends=True, lazy=lazy)
    if return_model_output and (minimum_frequency is not None or maximum_frequency is not None):
        # the reference zeroes these columns of the returned arrays (note_creation.py:338-341); all files share two arrays
        flo, fhi = infer.frequency_to_column_range(minimum_frequency, maximum_frequency)
        for k in ("note", "onset"):
            whole = outs[0][k].base if n and outs[0][k].base is not None else None
            for m in ([whole] if isinstance(whole, np.ndarray) and whole.ndim == 2 else [o[k] for o in outs]):
                m[:, :flo] = 0
                m[:, fhi:] = 0
    midis = [infer.LazyPrettyMIDI(ev, multiple_pitch_bends, midi_tempo) if build_midi else None for ev in events]
    if build_midi and not lazy:
        for m in midis:
            m.instruments  # assemble the Instrument / Note / PitchBend objects now
    return list(zip(outs, midis, events))


def predict_and_save(
    audio_path_list: Sequence[Union[pathlib.Path, str]],
    output_directory: Union[pathlib.Path, str],
    save_midi: bool,
    sonify_midi: bool,
    save_model_outputs: bool,
    save_notes: bool,
    model_or_model_path: Union[Model, str, pathlib.Path],
    onset_threshold: float = DEFAULT_ONSET_THRESHOLD,
    frame_threshold: float = DEFAULT_FRAME_THRESHOLD,
    minimum_note_length: float = DEFAULT_MINIMUM_NOTE_LENGTH_MS,
    minimum_frequency: Optional[float] = None,
    maximum_frequency: Optional[float] = None,
    multiple_pitch_bends: bool = False,
    melodia_trick: bool = True,
    debug_file: Optional[pathlib.Path] = None,
    sonification_samplerate: int = DEFAULT_SONIFICATION_SAMPLERATE,
    midi_tempo: float = DEFAULT_MINIMUM_MIDI_TEMPO,
) -> None:
    """reference: inference.py:509-604 — same files, names and failure behaviour (print, then re-raise).

    Several files without `debug_file` go through the batch path: one device pass per `BATCH_FILES` files
    (`predict_batch`: GPU ingest, one library call) and one `bp_write_note_files` call for their MIDI / CSV files
    (csrc/writers.cu) instead of a Python loop over files and notes."""

    def _saved(kind: str, path) -> None:
        print(f"  ✅ Saved {kind.lower().replace('_', ' ')} to {path}")

    def _failed(kind: str, path) -> None:
        print(f"\n🚨 Failed to save {kind.lower().replace('_', ' ')} to {path} \n")

    paths = list(audio_path_list)
    if len(paths) > 1 and debug_file is None:
        model = model_or_model_path if isinstance(model_or_model_path, Model) else default_model(model_or_model_path)
        BATCH_FILES = 64
        for c0 in range(0, len(paths), BATCH_FILES):
            chunk = paths[c0 : c0 + BATCH_FILES]
            for q in chunk:
                print(f"\nPredicting MIDI for {q}...")
            results = predict_batch([pathlib.Path(q) for q in chunk], model, onset_threshold, frame_threshold, minimum_note_length,
                                    minimum_frequency, maximum_frequency, multiple_pitch_bends, melodia_trick, midi_tempo)
            midi_paths: List[Optional[pathlib.Path]] = [None] * len(chunk)
            csv_paths: List[Optional[pathlib.Path]] = [None] * len(chunk)
            for i, (audio_path, (model_output, midi_data, _events)) in enumerate(zip(chunk, results)):
                if save_model_outputs:
                    path = build_output_path(audio_path, output_directory, OutputExtensions.MODEL_OUTPUT_NPZ)
                    try:
                        np.savez(path, basic_pitch_model_output=model_output)
                        _saved(OutputExtensions.MODEL_OUTPUT_NPZ.name, path)
                    except Exception:
                        _failed(OutputExtensions.MODEL_OUTPUT_NPZ.name, path)
                        raise
                if save_midi:
                    midi_paths[i] = build_output_path(audio_path, output_directory, OutputExtensions.MIDI)
                if sonify_midi:
                    path = build_output_path(audio_path, output_directory, OutputExtensions.MIDI_SONIFICATION)
                    try:
                        infer.sonify_midi(midi_data, path, sr=sonification_samplerate)
                        _saved(OutputExtensions.MIDI_SONIFICATION.name, path)
                    except Exception:
                        _failed(OutputExtensions.MIDI_SONIFICATION.name, path)
                        raise
                if save_notes:
                    csv_paths[i] = build_output_path(audio_path, output_directory, OutputExtensions.NOTE_EVENTS)
            if save_midi or save_notes:
                try:
                    infer.write_note_files([r[2] for r in results], midi_paths if save_midi else None,
                                           csv_paths if save_notes else None, multiple_pitch_bends, midi_tempo)
                except Exception:
                    for kind, plist in ((OutputExtensions.MIDI, midi_paths), (OutputExtensions.NOTE_EVENTS, csv_paths)):
                        for path in plist:
                            if path is not None and not path.exists():
                                _failed(kind.name, path)
                    raise
                for mp, cp in zip(midi_paths, csv_paths):
                    if mp is not None:
                        _saved(OutputExtensions.MIDI.name, mp)
                    if cp is not None:
                        _saved(OutputExtensions.NOTE_EVENTS.name, cp)
        return

    for audio_path in audio_path_list:
        print("")
        model_output, midi_data, note_events = predict(
            pathlib.Path(audio_path), model_or_model_path, onset_threshold, frame_threshold, minimum_note_length,
            minimum_frequency, maximum_frequency, multiple_pitch_bends, melodia_trick, debug_file, midi_tempo,
        )
        if save_model_outputs:
            path = build_output_path(audio_path, output_directory, OutputExtensions.MODEL_OUTPUT_NPZ)
            try:
                np.savez(path, basic_pitch_model_output=model_output)
                _saved(OutputExtensions.MODEL_OUTPUT_NPZ.name, path)
            except Exception:
                _failed(OutputExtensions.MODEL_OUTPUT_NPZ.name, path)
                raise
        if save_midi:
            path = build_output_path(audio_path, output_directory, OutputExtensions.MIDI)
            try:
                midi_data.write(str(path))
                _saved(OutputExtensions.MIDI.name, path)
            except Exception:
                _failed(OutputExtensions.MIDI.name, path)
                raise
        if sonify_midi:
            path = build_output_path(audio_path, output_directory, OutputExtensions.MIDI_SONIFICATION)
            try:
                infer.sonify_midi(midi_data, path, sr=sonification_samplerate)
                _saved(OutputExtensions.MIDI_SONIFICATION.name, path)
            except Exception:
                _failed(OutputExtensions.MIDI_SONIFICATION.name, path)
                raise
        if save_notes:
            path = build_output_path(audio_path, output_directory, OutputExtensions.NOTE_EVENTS)
            try:
                save_note_events(note_events, path)
                _saved(OutputExtensions.NOTE_EVENTS.name, path)
            except Exception:
                _failed(OutputExtensions.NOTE_EVENTS.name, path)
                raise
