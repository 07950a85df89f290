"""Numeric constants of the transcription path (values follow reference: basic_pitch/constants.py:21-47;
decode constants follow reference: basic_pitch/note_creation.py:40-49)."""
import numpy as np

FFT_HOP = 256
AUDIO_SAMPLE_RATE = 22050
AUDIO_N_CHANNELS = 1
AUDIO_WINDOW_LENGTH = 2  # seconds

SEMITONES_PER_OCTAVE = 12
NOTES_BINS_PER_SEMITONE = 1
CONTOURS_BINS_PER_SEMITONE = 3
ANNOTATIONS_BASE_FREQUENCY = 27.5  # A0
ANNOTATIONS_N_SEMITONES = 88
N_FREQ_BINS_NOTES = ANNOTATIONS_N_SEMITONES * NOTES_BINS_PER_SEMITONE  # 88
N_FREQ_BINS_CONTOURS = ANNOTATIONS_N_SEMITONES * CONTOURS_BINS_PER_SEMITONE  # 264

ANNOTATIONS_FPS = AUDIO_SAMPLE_RATE // FFT_HOP  # 86
ANNOTATION_HOP = 1.0 / ANNOTATIONS_FPS
ANNOT_N_FRAMES = ANNOTATIONS_FPS * AUDIO_WINDOW_LENGTH  # 172 frames per model window
AUDIO_N_SAMPLES = AUDIO_SAMPLE_RATE * AUDIO_WINDOW_LENGTH - FFT_HOP  # 43844 samples per model window

# windowing of long audio (reference: basic_pitch/inference.py:185-191, 302-305)
DEFAULT_OVERLAPPING_FRAMES = 30
OVERLAP_LEN = DEFAULT_OVERLAPPING_FRAMES * FFT_HOP  # 7680
HOP_SIZE = AUDIO_N_SAMPLES - OVERLAP_LEN  # 36164
FRAMES_PER_HOP = ANNOT_N_FRAMES - DEFAULT_OVERLAPPING_FRAMES  # 142

# decode
MIDI_OFFSET = 21
MAX_FREQ_IDX = 87
N_PITCH_BEND_TICKS = 8192
DEFAULT_MIN_NOTE_LEN = 11
ENERGY_TOLERANCE = 11
MAGIC_ALIGNMENT_OFFSET = 0.0018
MIDI_VELOCITY_SCALE = 127
PITCH_BEND_SCALE = 4096


# training-time sampling weights of the datasets (reference: constants.py:49-55; training is out of scope, kept so that
# `from basic_pitch.constants import *` users find the name)
DATASET_SAMPLING_FREQUENCY = {"MAESTRO": 5, "GuitarSet": 2, "MedleyDB-Pitch": 2, "iKala": 2, "slakh": 2}


def _freq_bins(bins_per_semitone: int, base_frequency: float, n_semitones: int) -> np.ndarray:
    step = 2.0 ** (1.0 / (SEMITONES_PER_OCTAVE * bins_per_semitone))
    return base_frequency * step ** np.arange(bins_per_semitone * n_semitones)


FREQ_BINS_NOTES = _freq_bins(NOTES_BINS_PER_SEMITONE, ANNOTATIONS_BASE_FREQUENCY, ANNOTATIONS_N_SEMITONES)
FREQ_BINS_CONTOURS = _freq_bins(CONTOURS_BINS_PER_SEMITONE, ANNOTATIONS_BASE_FREQUENCY, ANNOTATIONS_N_SEMITONES)
