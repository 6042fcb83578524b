"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol include/zero_b200.h
declares; the host-side config surface and plugin registry behave like the reference's (run.py, models/model.py)."""
import ctypes
import os
import re

import pytest

from zero_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "zero_b200.h")).read()
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        from zero_b200 import build
        build.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), "libzero_b200.so does not export %s" % sym
    assert sorted(L.EXPORTS) == declared
    lib.zb_abi_version.restype = ctypes.c_int
    assert lib.zb_abi_version() == 2


def test_ctypes_struct_sizes_match_header_layout():
    # spot-check: natural C alignment of the mirrored structs (pointers 8, int64 8, int32/float 4)
    assert ctypes.sizeof(L.GemmArgs) == 3 * 8 + 6 * 8 + 4 * 4 + 8 + 8 + 8 + 4 + 4
    assert ctypes.sizeof(L.AdamArgs) == 5 * 8 + 8 + 5 * 4 + 4 + 8 + 8


def test_mirrored_struct_layouts_match_the_library():
    lib = ctypes.CDLL(L.LIB_PATH)
    lib.zb_abi_struct_size.argtypes = [ctypes.c_int32]
    lib.zb_abi_struct_size.restype = ctypes.c_int64
    for which, cls in enumerate(L.STRUCTS):
        assert lib.zb_abi_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.zb_abi_struct_size(99) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.ZeroB200Error):
        L.load()


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zero_b200.engine import Engine
    from zero_b200.params import transformer_base
    with pytest.raises(L.ZeroB200Error):
        Engine(transformer_base(), 1000, 1000)
