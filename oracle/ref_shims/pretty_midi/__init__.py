"""Stand-in for the pretty_midi containers the reference decode fills (oracle-side only)."""


def instrument_name_to_program(name):
    assert name == "Electric Piano 1"
    return 4


class Note:
    def __init__(self, velocity, pitch, start, end):
        self.velocity, self.pitch, self.start, self.end = velocity, pitch, start, end


class PitchBend:
    def __init__(self, pitch, time):
        self.pitch, self.time = pitch, time


class Instrument:
    def __init__(self, program, is_drum=False, name=""):
        self.program, self.is_drum, self.name = program, is_drum, name
        self.notes, self.pitch_bends = [], []


class PrettyMIDI:
    def __init__(self, midi_file=None, resolution=220, initial_tempo=120.0):
        self.resolution, self.initial_tempo = resolution, initial_tempo
        self.instruments = []
