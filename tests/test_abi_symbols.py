"""The C-ABI library loads and exports every symbol include/pumas_b200.h
declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pumas_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pumas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"PUMAS_API[^;(]*?\b(pumas_\w+)\s*\(", text)
    assert len(names) > 40
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(pb.LIBRARY)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts():
    assert pb.STATE_DTYPE.itemsize == 104          # pumas.h:289-309
    assert ctypes.sizeof(pb.ContextStruct) == 96   # pumas.h:636-672
    assert ctypes.sizeof(pb.Layer) == 64
    assert ctypes.sizeof(pb.Geometry) == 16 + 16 + 72 + 8 * 64


def test_version_and_constants():
    lib = pb.library()
    major, minor, patch = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.pumas_version(ctypes.byref(major), ctypes.byref(minor),
                      ctypes.byref(patch))
    assert (major.value, minor.value, patch.value) == (1, 2, 3)
    v = ctypes.c_double()
    assert lib.pumas_constant(7, ctypes.byref(v)) == 0  # MUON_MASS
    assert v.value == 0.10565839
    assert lib.pumas_constant(99, ctypes.byref(v)) == 7
