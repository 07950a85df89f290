"""Alias of basic_pitch_b200.note_creation (drop-in for reference: basic_pitch/note_creation.py)."""
from basic_pitch_b200.note_creation import *  # noqa: F401,F403
from basic_pitch_b200 import note_creation as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
