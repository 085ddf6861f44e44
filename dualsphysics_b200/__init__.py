"""dualsphysics_b200 -- B200-native WCSPH time-step hot path behind the DualSPHysics case interface.

Only what the path needs lives here: ``csrc/`` (CUDA kernels, the C-ABI and the C++ host driver),
``lib.py`` (ctypes mirror of the interface), ``casegen.py`` / ``bi4.py`` (case writer and .bi4 i/o
tooling; GenCase is closed source) and ``b2sd.py`` (fixture container).
"""
from . import bi4, b2sd, casegen  # noqa: F401
from .lib import (Context, Simulation, Params, Info, Floating, B2SError, load_library, LIB_PATH)  # noqa: F401

__all__ = ["bi4", "b2sd", "casegen", "Context", "Simulation", "Params", "Info", "Floating", "B2SError",
           "load_library", "LIB_PATH"]
