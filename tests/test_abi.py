"""The C-ABI library loads and exports every symbol include/kmc_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from demo1 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_bound_symbols():
    assert declared_symbols() == sorted(nat.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(nat.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_record_layouts_match_header():
    assert nat.EVENT_COMPACT.itemsize == 16
    assert nat.EVENT_FULL.itemsize == 88
    assert nat.EVENT_FULL.fields["counts"][1] == 16


def test_no_device_is_a_loud_error():
    lib = nat.load()
    if lib.kmc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from demo1.engine import Engine
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Engine((8, 8), [1.0] * 17, [0])


def test_bad_dimensions_are_rejected_before_touching_the_device():
    lib = nat.load()
    h = C.c_void_p()
    rates = (C.c_double * 17)(*([1.0] * 17))
    order = (C.c_int * 1)(0)
    for dim in [(4, 8), (8, 6), (2, 2)]:
        assert lib.kmc_create(C.byref(h), 0, dim[0], dim[1], 1, rates, order, 1) == nat.KMC_ERR_ARG
        assert b"dimensions" in lib.kmc_last_error()
