"""gpuvmem_b200 — B200-native engine for gpuvmem's objective + gradient hot path.

The product is the C-ABI shared library ``libgvmb200.so`` (hand-written CUDA for
sm_100a, see ``include/gvm_b200.h``); this package is the thin Python binding used
by the tests, ``bench.py`` and the multi-GPU launcher, plus the synthetic
uv-coverage generator that stands in for Measurement-Set ingestion.

There is deliberately no CPU fallback: importing :mod:`gpuvmem_b200.lib` raises
if the library has not been built (``python -c 'import __graft_entry__ as g;
g.build()'``), and creating an engine raises without a CUDA device.
"""
from .lib import load_library, lib_path  # noqa: F401
from .engine import Engine, EngineError  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["Engine", "EngineError", "load_library", "lib_path", "synth"]
