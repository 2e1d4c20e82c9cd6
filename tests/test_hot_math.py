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


@pytest.mark.gpu
def test_division_and_square_root_are_ieee():
    """hot_div / hot_sqrt (the fast paths of div.rn.f64 / sqrt.rn.f64 without
    range check and slow path call) give the bits of a / b and sqrt(x): 2^28
    random pairs per exponent spread, one operand in four next to a power of
    two (mantissa 0 + j or all ones - j, j < 64: where the ties of the
    roundings live, e.g. 1 / (1 - 2^-53)), one numerator in 64 zero."""
    import pumas_b200 as pb
    assert pb.device_available()
    lib = pb.library()
    lib.pumas_b200_hot_math_selftest.argtypes = [
        ctypes.c_long, ctypes.c_ulong, ctypes.c_int,
        ctypes.POINTER(ctypes.c_ulonglong)]
    for spread in (1, 4, 60, 300, 480):
        bad = (ctypes.c_ulonglong * 2)()
        assert lib.pumas_b200_hot_math_selftest(
            1 << 28, 11 + spread, spread, bad) == 0
        assert (bad[0], bad[1]) == (0, 0), (spread, bad[0], bad[1])
