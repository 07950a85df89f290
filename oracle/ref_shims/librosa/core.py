import numpy as np


def frames_to_time(frames, sr=22050, hop_length=512):
    return (np.asanyarray(frames) * hop_length).astype(int) / float(sr)


def cqt_frequencies(n_bins, fmin, bins_per_octave=12):
    return fmin * 2.0 ** (np.arange(n_bins) / float(bins_per_octave))
