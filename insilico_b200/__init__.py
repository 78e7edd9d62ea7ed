"""insilico_b200 -- B200-native element-assembly engine (drop-in for inSilico's base/asmb hot path).

Layout: csrc/ (CUDA kernels + C ABI, built into lib/libinsilico_b200.so), engine.py (ctypes binding and the
host-side mirror of the reference interface), meshgen.py (synthetic meshes), partition.py (multi-GPU element
blocks).  Nothing in this package imports oracle/.
"""
from . import engine, meshgen  # noqa: F401
from .engine import Engine, EngineError  # noqa: F401
