"""Minimal MIDI containers + Standard MIDI File writer (used when `pretty_midi` is not installed).

The reference fills `pretty_midi.PrettyMIDI / Instrument / Note / PitchBend` objects and lets
pretty_midi/mido serialise them (reference: basic_pitch/note_creation.py:240-271,
basic_pitch/inference.py:574-584).  Those packages are optional here: these classes carry the same
attributes (`instruments`, `notes`, `pitch_bends`, `program`, `velocity`, `pitch`, `start`, `end`,
`time`) and `PrettyMIDI.write()` emits a format-1 file with the same event semantics as pretty_midi
(resolution 220 ticks per beat, one tempo, one track per instrument, channel = instrument index
skipping 9).  Byte-level equality with pretty_midi's output is not claimed (no golden .mid exists).
"""
from __future__ import annotations

import struct
from typing import List


def instrument_name_to_program(name: str) -> int:
    if name != "Electric Piano 1":
        raise ValueError("only 'Electric Piano 1' is known to this minimal table")
    return 4


class Note:
    def __init__(self, velocity: int, pitch: int, start: float, end: float):
        self.velocity, self.pitch, self.start, self.end = velocity, pitch, start, end

    def __repr__(self) -> str:
        return f"Note(start={self.start:f}, end={self.end:f}, pitch={self.pitch}, velocity={self.velocity})"


class PitchBend:
    def __init__(self, pitch: int, time: float):
        self.pitch, self.time = pitch, time

    def __repr__(self) -> str:
        return f"PitchBend(pitch={self.pitch:d}, time={self.time:f})"


class Instrument:
    def __init__(self, program: int, is_drum: bool = False, name: str = ""):
        self.program, self.is_drum, self.name = program, is_drum, name
        self.notes: List[Note] = []
        self.pitch_bends: List[PitchBend] = []
        self.control_changes: list = []


def _vlq(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


class PrettyMIDI:
    def __init__(self, midi_file=None, resolution: int = 220, initial_tempo: float = 120.0):
        if midi_file is not None:
            raise NotImplementedError("reading MIDI files is outside the hot path")
        self.resolution = resolution
        self.initial_tempo = initial_tempo
        self.instruments: List[Instrument] = []

    def time_to_tick(self, t: float) -> int:
        return int(round(t * self.resolution * self.initial_tempo / 60.0))

    def get_end_time(self) -> float:
        ends = [n.end for i in self.instruments for n in i.notes] + [b.time for i in self.instruments for b in i.pitch_bends]
        return max(ends) if ends else 0.0

    def synthesize(self, fs: int = 44100, wave=None):
        """Additive synthesis of the notes (what `pretty_midi.PrettyMIDI.synthesize` is used for by the reference's
        sonify_midi, note_creation.py:119-128): every note is `wave` (default np.sin) at its pitch, bent by the
        instrument's pitch-bend events (+-2 semitones full scale), scaled by velocity / 127 with 10 ms linear fades; the
        sum is normalised to a peak of 1.  A behavioural stand-in: not sample-identical to pretty_midi."""
        import numpy as np

        wave = np.sin if wave is None else wave
        end_time = self.get_end_time()
        out = np.zeros(int(fs * (end_time + 1)))
        if not self.instruments or out.size == 0:
            return np.array([]) if not self.instruments else out
        for inst in self.instruments:
            if inst.is_drum:
                continue
            bends = sorted(inst.pitch_bends, key=lambda b: b.time)
            bt = np.array([b.time for b in bends])
            bv = np.array([b.pitch for b in bends], dtype=np.float64) * (2.0 / 8192.0)  # semitones
            for n in inst.notes:
                a, b = int(fs * n.start), int(fs * n.end)
                if b <= a:
                    continue
                t = np.arange(a, b) / fs
                semis = np.full(b - a, float(n.pitch))
                if len(bt):  # zero-order hold of the latest bend event
                    k = np.searchsorted(bt, t, side="right") - 1
                    semis = semis + np.where(k >= 0, bv[np.maximum(k, 0)], 0.0)
                freq = 440.0 * 2.0 ** ((semis - 69.0) / 12.0)
                phase = 2.0 * np.pi * np.cumsum(freq) / fs
                env = np.ones(b - a)
                fade = min(int(0.01 * fs), (b - a) // 2)
                if fade > 0:
                    ramp = np.linspace(0.0, 1.0, fade, endpoint=False)
                    env[:fade] = ramp
                    env[-fade:] = ramp[::-1]
                out[a:b] += wave(phase) * env * (n.velocity / 127.0)
        peak = np.abs(out).max()
        return out / peak if peak > 0 else out

    def write(self, filename: str) -> None:
        tracks = []
        tempo_us = int(round(6e7 / self.initial_tempo))
        meta = b"\x00\xff\x51\x03" + struct.pack(">I", tempo_us)[1:] + b"\x00\xff\x58\x04\x04\x02\x18\x08" + b"\x01\xff\x2f\x00"
        tracks.append(meta)
        channels = [c for c in range(16) if c != 9]
        for idx, inst in enumerate(self.instruments):
            ch = 9 if inst.is_drum else channels[idx % len(channels)]
            ev = [(0, 0, bytes([0xC0 | ch, inst.program & 0x7F]))]
            for n in inst.notes:
                ev.append((self.time_to_tick(n.start), 2, bytes([0x90 | ch, n.pitch & 0x7F, max(0, min(127, n.velocity))])))
                ev.append((self.time_to_tick(n.end), 1, bytes([0x90 | ch, n.pitch & 0x7F, 0])))
            for b in inst.pitch_bends:
                v = int(b.pitch) + 8192
                ev.append((self.time_to_tick(b.time), 0, bytes([0xE0 | ch, v & 0x7F, (v >> 7) & 0x7F])))
            ev.sort(key=lambda e: (e[0], e[1]))
            data, last = bytearray(), 0
            for tick, _prio, msg in ev:
                data += _vlq(max(0, tick - last)) + msg
                last = max(last, tick)
            data += b"\x01\xff\x2f\x00"
            tracks.append(bytes(data))
        with open(filename, "wb") as fh:
            fh.write(b"MThd" + struct.pack(">IHHH", 6, 1, len(tracks), self.resolution))
            for t in tracks:
                fh.write(b"MTrk" + struct.pack(">I", len(t)) + t)
