"""The single precision log / exp restated in pb_dcs.h (glibc's algorithm,
see the comment there) against the system libm that the reference links:
dcs_evaluate (pumas.c:11279-11343) depends on their exact roundings."""
import ctypes

import numpy as np
import pytest

import harness


@pytest.fixture(scope="module")
def libs():
    if not harness.have_port():
        pytest.skip("oracle/_port not built")
    port = ctypes.CDLL(harness.PORT_SO)
    libm = ctypes.CDLL("libm.so.6")
    for lib, names in ((port, ("port_logf", "port_expf")),
                       (libm, ("logf", "expf"))):
        for name in names:
            f = getattr(lib, name)
            f.restype = ctypes.c_float
            f.argtypes = [ctypes.c_float]
    return port, libm


def test_logf_matches_libm(libs):
    port, libm = libs
    rng = np.random.default_rng(0)
    # arguments of dcs_evaluate: q / K in (1e-7, 1] and K / K0 in [1, 1.25]
    x = np.concatenate([
        np.exp(rng.uniform(np.log(1e-7), 0., 60000)),
        rng.uniform(1., 1.3, 30000),
        np.exp(rng.uniform(-80., 80., 10000))]).astype(np.float32)
    bad = sum(1 for v in x if port.port_logf(float(v)) != libm.logf(float(v)))
    assert bad == 0


def test_expf_matches_libm(libs):
    port, libm = libs
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-87., 88., 90000),
                        rng.uniform(-1., 1., 10000)]).astype(np.float32)
    bad = sum(1 for v in x if port.port_expf(float(v)) != libm.expf(float(v)))
    # glibc picks an FMA variant at run time on some CPUs: 2 arguments out of
    # 2.2e9 differ from the non-FMA evaluation (exhaustive scan)
    assert bad <= 1
