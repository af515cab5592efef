"""The C-ABI library loads and exports every symbol the headers declare (no compute: runs without a GPU)."""
import ctypes
import os
import re

import numpy as np

from slide_b200 import build, lib, program

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_builds_and_exports_declared_symbols():
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = set()
    for header in ("slide_b200.h", "slide_sap.h"):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(slide_[a-z0-9_]+)\s*\(", src))
    assert declared, "no declarations parsed"
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    for sym in declared:
        assert hasattr(handle, sym), sym
    assert handle.slide_abi_version() >= 1


def test_record_layout_matches_header():
    assert program.OP_DTYPE.itemsize == 8 + 8 * program.NPARAM + 4 * program.NFPARAM
    assert program.V["GEMM_NFIELD"] <= program.NPARAM and program.V["SM_NFIELD"] <= program.NPARAM
    assert program.V["GEMM_XFR"] == program.V["GEMM_XFA"] + program.V["XF_NFIELD"]
    b = program.Builder(2)
    t = b.tensor("a", 4, 6)
    assert t.ld == 8 and t.off % 256 == 0 and t.off >= b.stats_cap
    rec = b.pack()
    assert rec.dtype == program.OP_DTYPE and len(rec) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        lib.load()
    except lib.SlideError as e:
        assert "no CPU" in str(e)
    else:
        raise AssertionError("expected SlideError")
