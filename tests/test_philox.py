"""Philox4x32-10 stream (pumas_b200/csrc/philox.h): known-answer vectors of the
Random123 distribution (Salmon et al., SC'11, kat_vectors) and the stream
contract shared by the CPU oracle and the kernels."""
import ctypes

import numpy as np
import pytest

import harness

KAT = [
    ((0, 0, 0, 0), (0, 0),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2,
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
     (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.fixture(scope="module")
def lib():
    if not harness.have_port():
        pytest.skip("oracle/_port not built")
    lib = ctypes.CDLL(harness.PORT_SO)
    lib.port_philox_uniforms.argtypes = [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
    return lib


def test_known_answers(lib):
    for ctr, key, expected in KAT:
        c = (ctypes.c_uint32 * 4)(*ctr)
        k = (ctypes.c_uint32 * 2)(*key)
        out = (ctypes.c_uint32 * 4)()
        lib.port_philox_block(c, k, out)
        assert tuple(out) == expected


def test_stream_contract(lib):
    """Variate j of track t is word j % 4 of block j // 4, counter
    (block, 0, t_lo, t_hi), key (seed_lo, seed_hi), u = (w + 0.5) / 2^32."""
    seed, track, n = 0x1234567890abcdef, 0xfedcba9876543210, 37
    u = np.zeros(n)
    lib.port_philox_uniforms(seed, track, n, u.ctypes.data)
    key = (ctypes.c_uint32 * 2)(seed & 0xffffffff, seed >> 32)
    expected = []
    for block in range((n + 3) // 4):
        c = (ctypes.c_uint32 * 4)(block, 0, track & 0xffffffff, track >> 32)
        out = (ctypes.c_uint32 * 4)()
        lib.port_philox_block(c, key, out)
        expected += [(w + 0.5) / 4294967296. for w in out]
    assert np.array_equal(u, np.array(expected[:n]))
    assert (u > 0).all() and (u < 1).all()


def test_oracle_draws_the_same_stream(lib):
    if not harness.have_oracle():
        pytest.skip("oracle/_ref not built")
    ref = ctypes.CDLL(harness.ORACLE_SO)
    ref.oracle_philox_uniforms.argtypes = [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
    a, b = np.zeros(1001), np.zeros(1001)
    lib.port_philox_uniforms(7, 123456789, 1001, a.ctypes.data)
    ref.oracle_philox_uniforms(7, 123456789, 1001, b.ctypes.data)
    assert np.array_equal(a, b)
    assert abs(a.mean() - 0.5) < 0.05
