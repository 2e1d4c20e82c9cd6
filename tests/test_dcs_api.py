"""Host side DCS API of the C ABI (pumas_dcs_get / _default / _range,
pumas_physics_dcs, pumas_electronic_dcs; pumas.h:2110-2284) against the
reference itself (oracle/_ref exports the whole pumas.h API): the function
pointers returned evaluate the code the kernels run (pb_dcs.h,
pb_dcs_models.h) and must agree bit for bit with the reference's models."""
import ctypes

import numpy as np
import pytest

import harness
import pumas_b200 as pb

DCS_T = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_double, ctypes.c_double,
                         ctypes.c_double, ctypes.c_double, ctypes.c_double)
MODELS = {0: ("SSR", "KKP", "ABB"), 1: ("SSR", "KKP"),
          2: ("DRSS", "BM", "BBKS")}
# (Z, A) of H, C, O, Na, Fe, Pb, U and of the "Rk" standard rock pseudo element
TARGETS = ((1., 1.008), (6., 12.0108), (8., 15.9994), (11., 22.98977),
           (26., 55.845), (82., 207.2), (92., 238.02891), (11., 22.))
MASSES = (0.10565839, 1.77682)


def _prototype(lib):
    lib.pumas_dcs_get.argtypes = [ctypes.c_int, ctypes.c_char_p,
                                  ctypes.POINTER(ctypes.c_void_p)]
    lib.pumas_dcs_default.restype = ctypes.c_char_p
    lib.pumas_dcs_default.argtypes = [ctypes.c_int]
    lib.pumas_dcs_range.argtypes = [
        ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.pumas_physics_dcs.argtypes = [
        ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
        ctypes.POINTER(ctypes.c_void_p)]
    lib.pumas_electronic_dcs.restype = ctypes.c_double
    lib.pumas_electronic_dcs.argtypes = [ctypes.c_double] * 5
    lib.pumas_elastic_dcs.restype = ctypes.c_double
    lib.pumas_elastic_dcs.argtypes = [ctypes.c_double] * 5
    lib.pumas_elastic_path.restype = ctypes.c_double
    lib.pumas_elastic_path.argtypes = [ctypes.c_int] + [ctypes.c_double] * 4
    return lib


def _get(lib, process, model):
    f = ctypes.c_void_p()
    rc = lib.pumas_dcs_get(process, model.encode() if model else None,
                           ctypes.byref(f))
    assert rc == 0 and f.value
    return DCS_T(f.value)


@pytest.fixture(scope="module")
def libs(oracle):
    return _prototype(oracle.lib), _prototype(pb.library())


def test_models_bit_exact(libs):
    ref, ours = libs
    rng = np.random.default_rng(11)
    for process, names in MODELS.items():
        for name in names + (None,):
            f_ref, f_ours = _get(ref, process, name), _get(ours, process, name)
            for Z, A in TARGETS:
                for m in MASSES:
                    for _ in range(60):
                        k = float(np.exp(rng.uniform(np.log(1e-2),
                                                     np.log(1e8))))
                        q = float(k * np.exp(rng.uniform(np.log(1e-5), 0.)))
                        a, b = f_ref(Z, A, m, k, q), f_ours(Z, A, m, k, q)
                        assert a == b, (process, name, Z, A, m, k, q, a, b)


def test_default_range_electronic(libs):
    ref, ours = libs
    rng = np.random.default_rng(12)
    for process in range(3):
        assert ref.pumas_dcs_default(process) == \
            ours.pumas_dcs_default(process)
        for Z, _ in TARGETS:
            for m in MASSES:
                for k in np.exp(rng.uniform(np.log(1e-2), np.log(1e8), 20)):
                    a = (ctypes.c_double(), ctypes.c_double())
                    b = (ctypes.c_double(), ctypes.c_double())
                    assert ref.pumas_dcs_range(
                        process, Z, m, float(k), a[0], a[1]) == 0
                    assert ours.pumas_dcs_range(
                        process, Z, m, float(k), b[0], b[1]) == 0
                    assert (a[0].value, a[1].value) == \
                        (b[0].value, b[1].value), (process, Z, m, k)
    assert ours.pumas_dcs_default(3) is None
    # one bound only (NULL pointers allowed, pumas.c:14233-14238)
    b = ctypes.c_double()
    assert ours.pumas_dcs_range(1, 11., MASSES[0], 10., None, b) == 0
    for Z, I in ((1., 19.2e-9), (11., 136.4e-9), (82., 823e-9)):
        for m in MASSES:
            for _ in range(200):
                k = float(np.exp(rng.uniform(np.log(1e-3), np.log(1e8))))
                q = float(k * np.exp(rng.uniform(np.log(1e-7), 0.)))
                assert ref.pumas_electronic_dcs(Z, I, m, k, q) == \
                    ours.pumas_electronic_dcs(Z, I, m, k, q)


def test_elastic_dcs_bit_exact(libs):
    """pumas_elastic_dcs (pumas.c:15545): frame and screening parameters,
    spin and nuclear form factors of the hard elastic events (pb_coulomb.h)."""
    ref, ours = libs
    rng = np.random.default_rng(13)
    for Z, A in TARGETS:
        for m in MASSES:
            for _ in range(300):
                k = float(np.exp(rng.uniform(np.log(1e-4), np.log(1e7))))
                theta = float(np.exp(rng.uniform(np.log(1e-9), np.log(3.1))))
                a = ref.pumas_elastic_dcs(Z, A, m, k, theta)
                b = ours.pumas_elastic_dcs(Z, A, m, k, theta)
                # (both NaN beyond the kinematic limit of a light target)
                assert (a == b) or (a != a and b != b), (
                    Z, A, m, k, theta, a, b)
    # pumas_elastic_path (pumas.c:15584): pole reduction and transport
    # coefficients in long double on the host, as the reference
    for Z, A in TARGETS:
        for m in MASSES:
            for k in np.exp(rng.uniform(np.log(1e-4), np.log(1e7), 100)):
                for order in (0, 1):
                    a = ref.pumas_elastic_path(order, Z, A, m, float(k))
                    b = ours.pumas_elastic_path(order, Z, A, m, float(k))
                    assert (a == b) or (a != a and b != b), (
                        order, Z, A, m, k, a, b)
    assert ours.pumas_elastic_path(2, 11., 22., 0.1, 1.) == 0.
    assert ours.pumas_elastic_dcs(0.5, 1., 0.1, 1., 0.1) == 0.
    assert ours.pumas_elastic_dcs(1., 1., 0.1, 0., 0.1) == 0.


@pytest.mark.parametrize("dump, expected", [
    ("materials.muon", ("SSR", "SSR", "DRSS")),
    ("rock.muon.kkb", ("KKP", "KKP", "BBKS")),
    ("rock.muon.abm", ("ABB", "SSR", "BM"))])
def test_physics_models(libs, dump, expected):
    ref, ours = libs
    oracle = harness.Oracle(dump)
    physics = pb.Physics(harness.dump_path(dump))
    for process in range(3):
        names = []
        for lib, handle in ((ref, oracle.physics), (ours, physics._handle)):
            name, f = ctypes.c_char_p(), ctypes.c_void_p()
            assert lib.pumas_physics_dcs(
                handle, process, ctypes.byref(name), ctypes.byref(f)) == 0
            names.append(name.value.decode())
            g = ctypes.c_void_p()
            assert lib.pumas_dcs_get(process, name.value,
                                     ctypes.byref(g)) == 0
            assert f.value == g.value
        assert names[0] == names[1] == expected[process]


def test_errors():
    ours = _prototype(pb.library())
    f = ctypes.c_void_p(1)
    pb._last_error["rc"] = 0
    assert ours.pumas_dcs_get(3, b"SSR", ctypes.byref(f)) == 7  # INDEX_ERROR
    assert ours.pumas_dcs_get(0, b"XYZ", ctypes.byref(f)) == 14  # MODEL_ERROR
    assert f.value is None
    assert b"model XYZ not found for bremsstrahlung process" in \
        pb._last_error["message"].encode()
    a = ctypes.c_double()
    assert ours.pumas_dcs_range(-1, 1., 1., 1., a, a) == 7
    ours.pumas_dcs_register.argtypes = [ctypes.c_int, ctypes.c_char_p,
                                        ctypes.c_void_p]
    # pumas.c:15422-15460: nil function, nil name, bad process, known name
    assert ours.pumas_dcs_register(0, b"mine", None) == 24  # VALUE_ERROR
    fn = ctypes.c_void_p(1)
    assert ours.pumas_dcs_register(0, None, fn) == 24
    assert ours.pumas_dcs_register(3, b"mine", fn) == 7
    assert ours.pumas_dcs_register(0, b"SSR", fn) == 14  # MODEL_ERROR
    assert ours.pumas_physics_dcs(None, 0, None, None) == 8  # PHYSICS_ERROR
