"""Alias of basic_pitch_b200.constants (drop-in for reference: basic_pitch/constants.py)."""
from basic_pitch_b200.constants import *  # noqa: F401,F403
from basic_pitch_b200 import constants as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
