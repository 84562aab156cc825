"""The C-ABI library loads and exports every symbol include/gvm_b200.h declares; the
product has no CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gvm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gvm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gpuvmem_b200 import lib
    assert os.path.exists(lib.lib_path()), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    so = ctypes.CDLL(lib.lib_path())
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(so, n), f"{n} declared in include/gvm_b200.h but not exported"
    assert set(lib.SIGNATURES) == set(names), set(lib.SIGNATURES) ^ set(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpuvmem_b200 import Engine, EngineError
    with pytest.raises(EngineError, match="no CPU fallback|no CUDA device"):
        Engine(64, 64, -1e-5, 1e-5, 1e11)


def test_preprocessing_entry_points_refuse_to_run_without_a_gpu():
    """gvm_grid_reserve / gvm_weights (non-natural schemes) / gvm_grid_block fail loudly on a CPU-only box."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpuvmem_b200 import lib
    so = lib.load_library()
    assert so.gvm_grid_reserve(0, 64, 64, 1000, 1) != 0
    assert b"no CPU fallback" in so.gvm_last_error()


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gpuvmem_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "gvm_oracle" not in txt and "libgvref" not in txt, f
