"""Importable alias of the ``symbolic-music-diffusion_b200/`` package directory.

The task fixes the package directory name (which contains a hyphen and is therefore not a Python
identifier); this shim makes its modules importable as ``smd_b200.<module>``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "symbolic-music-diffusion_b200")
__path__.insert(0, _PKG_DIR)  # type: ignore[name-defined]

from .lib import load_library, SmdError  # noqa: E402,F401
from .engine import Engine, ModelConfig  # noqa: E402,F401
