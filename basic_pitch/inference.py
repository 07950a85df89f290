"""Alias of basic_pitch_b200.inference (drop-in for reference: basic_pitch/inference.py)."""
from basic_pitch_b200.inference import *  # noqa: F401,F403
from basic_pitch_b200 import inference as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
