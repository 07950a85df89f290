"""Stand-in for the few librosa calls the reference hot path makes (oracle-side only)."""
import numpy as np

from . import core  # noqa: F401


def midi_to_hz(notes):
    return 440.0 * (2.0 ** ((np.asanyarray(notes) - 69.0) / 12.0))


def hz_to_midi(frequencies):
    return 12 * (np.log2(np.asanyarray(frequencies)) - np.log2(440.0)) + 69


def load(path, sr=22050, mono=True):
    """WAV only. Other sample rates go through a Kaiser-windowed polyphase FIR (the reference uses
    librosa's soxr_hq; that library is not available, see tests/golden/README.md)."""
    from scipy.io import wavfile
    import scipy.signal

    fs, x = wavfile.read(path)
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(2 ** (8 * x.dtype.itemsize - 1))
    elif x.dtype.kind == "u":
        x = (x.astype(np.float32) - 128.0) / 128.0
    x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1) if mono else x.T
    if fs != sr:
        from math import gcd

        g = gcd(int(fs), int(sr))
        up, down = int(sr) // g, int(fs) // g
        cutoff = 0.913 / max(up, down)
        width = (1.0 - 0.913) / max(up, down)
        numtaps, beta = scipy.signal.kaiserord(125.0, width)
        numtaps |= 1
        h = scipy.signal.firwin(numtaps, cutoff + width / 2, window=("kaiser", beta))
        x = scipy.signal.resample_poly(x.astype(np.float64), up, down, window=h).astype(np.float32)
    return x, sr
