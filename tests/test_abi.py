"""The C-ABI library loads and exports every symbol include/gpyreg_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "gpyreg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from gpyreg_b200 import _build, _lib
    _build.build_library()          # no-op when the .so is newer than the sources
    return _lib.load()


def test_header_matches_binding(lib):
    from gpyreg_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.SYMBOLS) == syms


def test_every_symbol_exported(lib):
    raw = ctypes.CDLL(os.path.join(ROOT, "gpyreg_b200", "libgpyreg_b200.so"))
    for name in header_symbols():
        assert hasattr(raw, name), name
    assert lib.gpb_version() >= 100


def test_no_gpu_fails_loudly(lib):
    """Without a device the engine must raise, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpyreg_b200 import Engine, GpbError
    with pytest.raises(GpbError):
        Engine(0)
    assert b"CUDA" in lib.gpb_last_error(None) or b"device" in lib.gpb_last_error(None)
