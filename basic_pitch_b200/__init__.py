"""Import alias: the package directory is `basic-pitch_b200/` (not importable by name because of the
hyphen); this stub makes `import basic_pitch_b200` resolve to it."""
import os as _os

_real = _os.path.abspath(_os.path.join(_os.path.dirname(__file__), "..", "basic-pitch_b200"))
if not _os.path.isfile(_os.path.join(_real, "__init__.py")):
    raise ImportError(f"basic_pitch_b200: package directory {_real} not found")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "rb") as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
del _fh, _real
