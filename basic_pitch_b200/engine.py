"""Batch / multi-GPU host plumbing around the C ABI (PyTorch is used for device memory, pinned host
memory, streams and torch.distributed only — never for arithmetic).

  * `shard_range`          static contiguous partition of files (or windows) over ranks — the path has
                           no exchange step, so this is all the "parallelism" there is
  * `broadcast_weights`    the single collective: rank 0's parameter block -> every rank over NCCL
  * `PackedAudio`          a batch of clips back to back in pinned host memory (+ offsets)
  * `transcribe_packed`    audio (pinned host or device tensor) -> note events through ONE library call
  * `forward_device`       (B,43844) CUDA tensor -> posteriorgram CUDA tensors on the current stream
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .constants import AUDIO_N_SAMPLES, N_FREQ_BINS_CONTOURS, N_FREQ_BINS_NOTES
from .inference import Model


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_items` owned by `rank` (sizes differ by at most one)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class _CudaBlock:
    """Raw device memory exposed through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None,
        }  # fmt: skip


def param_block_tensor(model: Model):
    """The model's packed parameters (float32, on its device) as a torch tensor sharing memory."""
    import torch

    ptr, n = C.c_void_p(), C.c_size_t()
    model._lib.bp_model_param_block(model.handle, C.byref(ptr), C.byref(n))
    return torch.as_tensor(_CudaBlock(ptr.value, n.value), device=f"cuda:{model.device}")


def broadcast_weights(model: Model, src: int = 0, group=None) -> None:
    """One broadcast of the ~140 KB parameter block at init (NCCL on GPUs; any backend that can move the
    tensor works), then re-derive the kernel-side weight layouts."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    t = param_block_tensor(model)
    if dist.get_backend(group) == "nccl":
        dist.broadcast(t, src=src, group=group)
    else:  # CPU backends (tests): stage through host memory
        h = t.cpu()
        dist.broadcast(h, src=src, group=group)
        t.copy_(h)
    torch.cuda.synchronize(model.device)
    model._lib.bp_model_refresh(model.handle)


class PackedAudio:
    """Clips laid back to back in one float32 buffer with int64 offsets (the layout bp_transcribe_* takes)."""

    def __init__(self, clips: Sequence[np.ndarray], pinned: bool = True):
        import torch

        self.offsets = np.zeros(len(clips) + 1, dtype=np.int64)
        for i, c in enumerate(clips):
            self.offsets[i + 1] = self.offsets[i] + len(c)
        n = max(int(self.offsets[-1]), 1)
        self.tensor = torch.empty(n, dtype=torch.float32, pin_memory=pinned and torch.cuda.is_available())
        self.array = self.tensor.numpy()
        for i, c in enumerate(clips):
            self.array[self.offsets[i] : self.offsets[i + 1]] = c
        self.n_files = len(clips)

    @property
    def nbytes(self) -> int:
        return int(self.offsets[-1]) * 4

    def to_device(self, device: int):
        return self.tensor.to(f"cuda:{device}", non_blocking=True)


class NoteBuffers:
    """Reusable pinned output arrays for note events."""

    def __init__(self, n_files: int, note_cap: int, bend_cap: int):
        import torch

        pin = torch.cuda.is_available()

        def mk(n, dt):
            return torch.empty(max(n, 1), dtype=dt, pin_memory=pin)

        self.t = {
            "note_off": mk(n_files + 1, torch.int32), "start": mk(note_cap, torch.int32), "end": mk(note_cap, torch.int32),
            "pitch": mk(note_cap, torch.int32), "amp": mk(note_cap, torch.float32), "bend_off": mk(note_cap + 1, torch.int32),
            "bends": mk(bend_cap, torch.int32), "frame_off": mk(n_files + 1, torch.int64),
        }  # fmt: skip
        self.a = {k: v.numpy() for k, v in self.t.items()}
        self.notes = _lib.Notes()
        self.notes.note_capacity, self.notes.bend_capacity = note_cap, bend_cap
        a = self.a
        self.notes.note_off, self.notes.start_frame, self.notes.end_frame = a["note_off"].ctypes.data, a["start"].ctypes.data, a["end"].ctypes.data
        self.notes.pitch_midi, self.notes.amplitude = a["pitch"].ctypes.data, a["amp"].ctypes.data
        self.notes.bend_off, self.notes.bends = a["bend_off"].ctypes.data, a["bends"].ctypes.data
        self.n_files = n_files

    def result_bytes(self) -> int:
        n = int(self.a["note_off"][self.n_files])
        nb = int(self.a["bend_off"][n])
        return 4 * (self.n_files + 1) + n * 16 + 4 * (n + 1) + 4 * nb

    def n_notes(self) -> int:
        return int(self.a["note_off"][self.n_files])


def default_params(model: Model, **kw) -> _lib.DecodeParams:
    p = _lib.DecodeParams()
    model._lib.bp_default_decode_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def transcribe_packed_host(model: Model, audio: PackedAudio, out: NoteBuffers, params: Optional[_lib.DecodeParams] = None) -> int:
    """HOST buffers in, note events out: H2D of the audio and D2H of the events happen inside the call
    (bp_transcribe_host).  Returns the number of notes."""
    p = params or default_params(model)
    model._lib.bp_transcribe_host(model.handle, audio.array.ctypes.data, audio.offsets.ctypes.data, audio.n_files,
                                  C.byref(p), None, None, None, out.a["frame_off"].ctypes.data, C.byref(out.notes))
    return out.n_notes()


def transcribe_packed_device(model: Model, d_audio, offsets: np.ndarray, out: NoteBuffers,
                             params: Optional[_lib.DecodeParams] = None, stream: Optional[int] = None) -> int:
    """Audio already resident in HBM (a CUDA float32 tensor): bp_transcribe_device on torch's current stream."""
    import torch

    p = params or default_params(model)
    st = torch.cuda.current_stream(model.device).cuda_stream if stream is None else stream
    model._lib.bp_transcribe_device(model.handle, d_audio.data_ptr(), offsets.ctypes.data, len(offsets) - 1,
                                    C.byref(p), out.a["frame_off"].ctypes.data, C.byref(out.notes), st)
    return out.n_notes()


def forward_device(model: Model, d_audio, stream: Optional[int] = None):
    """(B, 43844) CUDA float32 tensor -> (note, onset, contour) CUDA tensors; asynchronous on the stream."""
    import torch

    if d_audio.dim() != 2 or d_audio.shape[1] != AUDIO_N_SAMPLES or d_audio.dtype != torch.float32 or not d_audio.is_contiguous():
        raise ValueError("expected a contiguous (B, 43844) float32 CUDA tensor")
    n = d_audio.shape[0]
    dev = d_audio.device
    note = torch.empty((n, 172, N_FREQ_BINS_NOTES), dtype=torch.float32, device=dev)
    onset = torch.empty((n, 172, N_FREQ_BINS_NOTES), dtype=torch.float32, device=dev)
    contour = torch.empty((n, 172, N_FREQ_BINS_CONTOURS), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
    model._lib.bp_forward_device(model.handle, d_audio.data_ptr(), n, note.data_ptr(), onset.data_ptr(), contour.data_ptr(), st)
    return note, onset, contour


def split_results(out: NoteBuffers) -> List[Dict[str, np.ndarray]]:
    return Model._split_notes({k: out.a[k] for k in ("note_off", "start", "end", "pitch", "amp", "bend_off", "bends")}, out.n_files)
