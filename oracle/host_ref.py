"""ORACLE (test infrastructure): windowing / unwrap arithmetic of the reference inference loop.

Restates reference: basic_pitch/inference.py:194-219 (`window_audio_file`), :222-244
(`get_audio_input`, minus file decoding), :247-279 (`unwrap_output`), :302-305 (overlap constants).
"""
import numpy as np

N_SAMPLES = 43844
N_OVERLAP_FRAMES = 30
OVERLAP_LEN = N_OVERLAP_FRAMES * 256  # 7680
HOP_SIZE = N_SAMPLES - OVERLAP_LEN  # 36164


def window_audio(audio: np.ndarray) -> np.ndarray:
    """mono float32 (n,) -> (n_windows, 43844): prepend overlap/2 zeros, hop 36164, zero-pad the tail."""
    x = np.concatenate([np.zeros(OVERLAP_LEN // 2, dtype=np.float32), audio.astype(np.float32)])
    starts = range(0, x.shape[0], HOP_SIZE)
    out = np.zeros((len(starts), N_SAMPLES), dtype=np.float32)
    for w, s in enumerate(starts):
        seg = x[s : s + N_SAMPLES]
        out[w, : len(seg)] = seg
    return out


def unwrap(output: np.ndarray, original_length: int) -> np.ndarray:
    """(n_windows,172,F) -> (T,F): drop 15 frames each side, concatenate, trim."""
    half = N_OVERLAP_FRAMES // 2
    o = output[:, half:-half, :]
    flat = o.reshape(o.shape[0] * o.shape[1], o.shape[2])
    n_keep = int(original_length / HOP_SIZE * (172 - N_OVERLAP_FRAMES))
    return flat[:n_keep, :]
