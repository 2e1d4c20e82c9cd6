"""The elementary functions of the lean step on the GPU (pb_hot_math.h),
checked on the CPU through test hooks of the CPU build of the core: the table
driven logarithm of a uniform variate (w + 1/2) 2^-32 against the host libm
and against a 60 digit reference (mpmath-free: Python's decimal module)."""
import ctypes
import math
from decimal import Decimal, getcontext

import numpy as np
import pytest

import harness


def _lib():
    if not harness.have_port():
        pytest.skip("oracle/_port not built")
    lib = ctypes.CDLL(harness.PORT_SO)
    lib.port_hot_log_uniform.restype = ctypes.c_double
    lib.port_hot_log_uniform.argtypes = [ctypes.c_uint32]
    lib.port_hot_log_uniform_n.argtypes = [
        ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
    return lib


def test_log_uniform_against_libm():
    """<= 1 ulp from glibc's log over 2 M words spread over the whole range,
    including the ends (u -> 0 and u -> 1) and every power of two."""
    lib = _lib()
    rng = np.random.default_rng(1)
    w = np.concatenate([
        rng.integers(0, 1 << 32, 2_000_000, dtype=np.uint64),
        np.arange(0, 50000, dtype=np.uint64),
        (1 << 32) - 1 - np.arange(0, 50000, dtype=np.uint64),
        np.concatenate([(1 << b) + np.arange(-200, 200) for b in range(9, 32)]
                       ).astype(np.uint64),
    ]).astype(np.uint32)
    got = np.empty(len(w))
    lib.port_hot_log_uniform_n(w.ctypes.data, len(w), got.ctypes.data)
    u = (w.astype(np.float64) + 0.5) * 2.0 ** -32
    ref = np.log(u)
    ulp = np.abs(np.nextafter(ref, 0.) - ref)
    err = np.abs(got - ref) / ulp
    assert err.max() <= 1.0, (err.max(), int(w[err.argmax()]))
    assert (got < 0.).all()


def test_log_uniform_against_exact():
    """<= 1.5 ulp from the exact value (60 digits) on a sample."""
    lib = _lib()
    getcontext().prec = 60
    rng = np.random.default_rng(2)
    words = [0, 1, 2, (1 << 32) - 1, (1 << 32) - 2, 1 << 31, (1 << 31) - 1]
    words += [int(x) for x in rng.integers(0, 1 << 32, 3000, dtype=np.uint64)]
    worst = 0.
    for w in words:
        exact = (Decimal(2 * w + 1) / Decimal(1 << 33)).ln()
        got = lib.port_hot_log_uniform(w)
        ulp = math.ulp(float(exact))
        worst = max(worst, abs(float(Decimal(got) - exact)) / ulp)
    assert worst <= 1.5, worst
