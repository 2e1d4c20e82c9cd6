"""CPU parity of the restated algorithm: the transport core of the kernels
(pumas_b200/csrc/pb_*.h) compiled for the host (oracle/_port) against the
reference itself (oracle/_ref), track by track with the same Philox streams.

With the same libm and no FMA contraction on either side the restatement is
expected to be BIT EXACT, including the number of variates each track draws.
The general instantiation, the compile-time specialised ones and the lean
step of pb_hot.h (mode 2: what the GPU library runs for the straggled mode in a
uniform medium, through the same slot packing) are covered, and the double-double replacement of
the reference's x87 long double arithmetic (pb_dd.h)."""
import os

import numpy as np
import pytest

import harness
from workloads import WORKLOADS

N = 120
THREADS = min(8, os.cpu_count() or 1)

_cache = {}


def sides(dump):
    if not (harness.have_oracle() and harness.have_port()):
        pytest.skip("oracle/_ref or oracle/_port not built")
    if dump not in _cache:
        _cache[dump] = (harness.Oracle(dump), harness.Port(dump),
                        harness.Port(dump, dd=True))
    return _cache[dump]


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_bit_exact(name):
    dump, builder, kwargs = WORKLOADS[name]
    oracle, port, port_dd = sides(dump)
    w = builder(oracle.material, N, **kwargs)
    so, eo, mo, ro, rc, _ = oracle.transport(
        w["settings"], w["geometry"], w["states"], seed=3, offset=17,
        threads=THREADS)
    assert rc == 0
    for engine, specialise in ((port, 1), (port, 0), (port, 2), (port_dd, 1)):
        engine.lib.port_set_specialise(specialise)
        sp, ep, mp, rp, steps, rcp = engine.transport(
            w["settings"], w["geometry"], w["states"], seed=3, offset=17,
            threads=THREADS)
        assert rcp == 0
        c = harness.compare_states(so, sp, eo, ep, ro, rp)
        assert c["n_diverged"] == 0
        assert c["n_bitexact"] == N, (name, specialise, c["max_rel"])
        assert np.array_equal(mo, mp)
    port.lib.port_set_specialise(1)


@pytest.mark.parametrize("dump, n_elements", [
    ("materials.muon", 6), ("rock.muon.kkb", 3), ("rock.muon.abm", 3)])
def test_dcs_models_bit_exact(dump, n_elements):
    """dcs_evaluate and the exact DCS model of every process, for the default
    models (SSR / SSR / DRSS) and the others (KKP / KKP / BBKS and
    ABB / SSR / BM, pb_dcs_models.h)."""
    oracle, port, port_dd = sides(dump)
    rng = np.random.default_rng(7)
    for process in range(4):
        for element in range(n_elements):
            for _ in range(300):
                k = float(np.exp(rng.uniform(np.log(1e-2), np.log(1e8))))
                q = float(k * np.exp(rng.uniform(np.log(1e-4), 0.)))
                for f in ("dcs_evaluate", "dcs_model"):
                    a = getattr(oracle.lib, "oracle_" + f)(
                        oracle.physics, process, element, k, q)
                    b = getattr(port.lib, "port_" + f)(
                        port.physics, process, element, k, q)
                    assert a == b, (f, process, element, k, q, a, b)


def test_unit_functions_bit_exact():
    """dcs_evaluate, the exact DCS models, msc and hard-elastic lookups."""
    oracle, port, port_dd = sides("materials.muon")
    rng = np.random.default_rng(5)
    import ctypes
    for process in range(4):
        for element in range(6):
            for _ in range(150):
                k = float(np.exp(rng.uniform(np.log(1e-2), np.log(1e7))))
                q = float(k * np.exp(rng.uniform(np.log(1e-3), 0.)))
                for f in ("dcs_evaluate", "dcs_model"):
                    a = getattr(oracle.lib, "oracle_" + f)(
                        oracle.physics, process, element, k, q)
                    b = getattr(port.lib, "port_" + f)(
                        port.physics, process, element, k, q)
                    assert a == b, (f, process, element, k, q)
    for material in range(4):
        for k in np.exp(rng.uniform(np.log(2e-3), np.log(1e8), 200)):
            k = float(k)
            assert oracle.lib.oracle_coulomb_ehs_length(
                oracle.physics, material, k) == \
                port.lib.port_coulomb_ehs_length(port.physics, material, k)
            for scheme in (-1, 0, 1, 2):
                a = (ctypes.c_double(), ctypes.c_double())
                b = (ctypes.c_double(), ctypes.c_double())
                oracle.lib.oracle_table_get_msc(
                    oracle.physics, scheme, material, k, a[0], a[1])
                port.lib.port_table_get_msc(
                    port.physics, scheme, material, k, b[0], b[1])
                assert (a[0].value, a[1].value) == (b[0].value, b[1].value)


def test_hard_elastic_cross_sections_double_double():
    """Element choice in a hard elastic event (pumas.c:5786-5826): the
    reference sums the pole reduction in x87 long double, the GPU in
    double-double (pb_dd.h). Measured against IEEE binary128 (libquadmath build
    of the same code), double-double is exact to the last double bit while the
    reference's long double loses up to ~2e-5 to cancellation at high energy:
    the GPU is the more accurate of the two, and the two can pick a different
    element with a probability of that order per hard elastic event in a
    compound material. The long double CPU build reproduces the reference bit
    for bit."""
    oracle, port, port_dd = sides("materials.muon")
    import ctypes
    quad_so = os.path.join(os.path.dirname(harness.PORT_SO),
                           "libpumas_port_quad.so")
    quad = None
    if os.path.exists(quad_so):
        quad = harness.Port("materials.muon")
        quad.lib = ctypes.CDLL(quad_so)
        quad.lib.port_physics_load.restype = ctypes.c_void_p
        quad.lib.port_physics_load.argtypes = [ctypes.c_char_p]
        quad.lib.port_coulomb_cs_hard.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
            ctypes.POINTER(ctypes.c_double)]
        quad.physics = quad.lib.port_physics_load(
            harness.dump_path("materials.muon").encode())
    rng = np.random.default_rng(9)
    worst_x87, worst_dd = 0., 0.
    for material in (1, 2, 3):  # Water, Air, WetRock: several elements
        for k in np.exp(rng.uniform(np.log(1e-3), np.log(1e6), 120)):
            a = (ctypes.c_double * 8)()
            b = (ctypes.c_double * 8)()
            c = (ctypes.c_double * 8)()
            t = (ctypes.c_double * 8)()
            n = oracle.lib.oracle_coulomb_cs_hard(
                oracle.physics, material, 2, float(k), a)
            port.lib.port_coulomb_cs_hard(port.physics, material, 2, float(k), b)
            port_dd.lib.port_coulomb_cs_hard(
                port_dd.physics, material, 2, float(k), c)
            if quad is not None:
                quad.lib.port_coulomb_cs_hard(
                    quad.physics, material, 2, float(k), t)
            for i in range(n):
                assert a[i] == b[i]  # long double build: identical
                if a[i] != 0.:
                    worst_x87 = max(worst_x87, abs(c[i] - a[i]) / abs(a[i]))
                if (quad is not None) and (t[i] != 0.):
                    worst_dd = max(worst_dd, abs(c[i] - t[i]) / abs(t[i]))
    assert worst_x87 < 1e-4, worst_x87
    if quad is not None:
        assert worst_dd < 1e-15, worst_dd


def test_user_media_callbacks_bit_exact():
    """The callback interface of the engine (CallbackGeometry, pb_media.h) on
    the CPU build with host functions -- the spherical shells of
    examples/user_media.h -- against the reference driving the same header
    through its own host callbacks: bit for bit (same libm on both sides).
    The GPU runs the same engine code with `__device__` functions
    (tests/test_callbacks.py)."""
    import ctypes
    import _user_media_run as um
    import pumas_b200 as pb
    from workloads import isotropic, log_uniform
    oracle, port, _ = sides("materials.muon")
    g, materials = um.shells(oracle.material)
    oracle.lib.oracle_set_user_media.argtypes = [
        ctypes.POINTER(um.UserMedia), ctypes.POINTER(ctypes.c_int)]
    f = port.lib.port_transport_user_media
    f.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(um.UserMedia),
        ctypes.POINTER(ctypes.c_int), ctypes.c_uint64, ctypes.c_uint64,
        ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]
    n = 3000
    cases = [
        (harness.make_settings(pb.MODE_STRAGGLED, pb.MODE_MIXED),
         pb.make_states(n, log_uniform(n, 0.5, 200., 1),
                        direction=isotropic(n, 2), position=(10., 5., -20.))),
        (harness.make_settings(pb.MODE_STRAGGLED, pb.MODE_MIXED,
                               direction=pb.MODE_BACKWARD,
                               event=pb.EVENT_LIMIT_ENERGY, limit_energy=1e3),
         pb.make_states(n, log_uniform(n, 0.5, 50., 3),
                        direction=isotropic(n, 4), position=(10., 5., -20.))),
        (harness.make_settings(pb.MODE_CSDA, pb.MODE_DISABLED),
         pb.make_states(n, log_uniform(n, 0.5, 200., 5),
                        direction=isotropic(n, 6), position=(0., 60., 0.))),
        (harness.make_settings(pb.MODE_MIXED, pb.MODE_MIXED,
                               event=pb.EVENT_MEDIUM),
         pb.make_states(n, log_uniform(n, 5., 500., 7),
                        direction=isotropic(n, 8), position=(0., 0., 45.))),
    ]
    try:
        oracle.lib.oracle_set_user_media(ctypes.byref(g), materials)
        for settings, states in cases:
            so, eo, mo, ro, rc, _ = oracle.transport(
                settings, pb.uniform_geometry(0), states, seed=11, offset=500,
                threads=THREADS)
            assert rc == 0
            sp = np.ascontiguousarray(states.copy())
            ep = np.zeros(n, dtype=np.int32)
            mp = np.zeros((n, 2), dtype=np.int32)
            rp = np.zeros(n, dtype=np.uint32)
            steps = np.zeros(n, dtype=np.uint32)
            assert f(port.physics, ctypes.byref(settings), ctypes.byref(g),
                     materials, 11, 500, n, sp.ctypes.data, ep.ctypes.data,
                     mp.ctypes.data, rp.ctypes.data, steps.ctypes.data) == 0
            c = harness.compare_states(so, sp, eo, ep, ro, rp)
            assert c["n_diverged"] == 0
            assert c["n_bitexact"] == n, c["max_rel"]
            assert np.array_equal(mo, mp)
            assert ((ep & pb.EVENT_MEDIUM) != 0).any() or (mp[:, 1] < 0).any()
    finally:
        oracle.lib.oracle_set_user_media(None, None)


def test_user_dcs_bit_exact():
    """A user registered DCS model (examples/user_dcs.h; tables made by the
    reference with it) evaluated by the engine through Physics::user_dcs, CPU
    build against the reference: bit for bit."""
    import ctypes
    from workloads import dcs_models
    import pumas_b200 as pb
    if not (harness.have_oracle() and harness.have_port()):
        pytest.skip("oracle/_ref or oracle/_port not built")
    lib_o = ctypes.CDLL(harness.ORACLE_SO)
    assert lib_o.oracle_register_user_dcs() == 0
    oracle = harness.Oracle("rock.muon.usr")
    port = harness.Port("rock.muon.usr")
    port.lib.port_register_user_dcs.argtypes = [ctypes.c_void_p]
    assert port.lib.port_register_user_dcs(port.physics) == 1
    # the model itself, then the transport that calls it
    K = np.logspace(-1, 6, 29)
    for k in K:
        for x in (1e-3, 0.05, 0.5, 0.9):
            a = oracle.lib.oracle_dcs_model(oracle.physics, 0, 0, k, x * k)
            b = port.lib.port_dcs_model(port.physics, 0, 0, k, x * k)
            assert a == b, (k, x, a, b)
    n = 2000
    for direction in (pb.MODE_FORWARD, pb.MODE_BACKWARD):
        w = dcs_models(oracle.material, n, direction=direction)
        so, eo, mo, ro, rc, _ = oracle.transport(
            w["settings"], w["geometry"], w["states"], seed=13, offset=900,
            threads=THREADS)
        assert rc == 0
        for specialise in (1, 0):
            port.lib.port_set_specialise(specialise)
            sp, ep, mp, rp, steps, rcp = port.transport(
                w["settings"], w["geometry"], w["states"], seed=13,
                offset=900, threads=THREADS)
            assert rcp == 0
            c = harness.compare_states(so, sp, eo, ep, ro, rp)
            assert c["n_diverged"] == 0
            assert c["n_bitexact"] == n, c["max_rel"]
    port.lib.port_set_specialise(1)
