"""The host side of the C ABI (libpumas_b200.so): dump loader and property
getters against the reference (oracle/_ref). No GPU needed: these entry points
are host functions in the reference too (pumas.c:2049, 2854-3111)."""
import ctypes
import filecmp
import os

import numpy as np
import pytest

import harness
import pumas_b200 as pb
from harness import MODE_CSDA, MODE_DISABLED, MODE_MIXED


@pytest.fixture(scope="module")
def physics():
    return pb.Physics(harness.dump_path())


def test_dump_round_trip(physics, tmp_path):
    out = tmp_path / "again.pumas"
    physics.dump(str(out))
    assert filecmp.cmp(str(out), harness.dump_path(), shallow=False)


def test_particle_and_materials(physics, oracle):
    kind, ctau, mass = physics.particle()
    assert (kind, ctau, mass) == (0, 658.654, 0.10565839)
    for name in ("StandardRock", "Water", "Air", "WetRock"):
        assert physics.material_index(name) == oracle.material(name)
    with pytest.raises(pb.PumasError) as e:
        physics.material_index("Unobtainium")
    assert e.value.rc == 22  # PUMAS_RETURN_UNKNOWN_MATERIAL


def test_tau_dump():
    tau = pb.Physics(harness.dump_path("ice.tau"))
    kind, ctau, mass = tau.particle()
    assert kind == 1 and ctau == 87.03e-6 and mass == 1.77682
    assert tau.material_index("Ice") == 0


def test_properties_bit_exact(physics, oracle):
    """Every property getter, all materials, energies across and beyond the
    grid: identical doubles."""
    energies = np.concatenate([
        np.exp(np.linspace(np.log(5e-4), np.log(5e9), 150)), [0., 1e-3, 1e9]])
    cases = []
    for m in range(4):
        for k in energies:
            k = float(k)
            for mode in (MODE_CSDA, MODE_MIXED):
                cases += [("range", mode, m, k), ("proper_time", mode, m, k),
                          ("stopping_power", mode, m, k),
                          ("kinetic_energy", mode, m, k * 1e3)]
            cases += [("magnetic_rotation", m, k), ("energy_straggling", m, k),
                      ("cross_section", m, k)]
            if k > 0.:
                cases += [("elastic_cutoff_angle", m, k), ("elastic_path", m, k)]
                for mode in (MODE_DISABLED, MODE_CSDA, MODE_MIXED):
                    cases.append(("transport_path", mode, m, k))
    for name, *args in cases:
        rc, expected = oracle.property(name, *args)
        assert rc == 0
        got = physics.property(name, *args)
        assert got == expected or (np.isnan(got) and np.isnan(expected)), (
            name, args, got, expected)


def test_property_errors(physics):
    with pytest.raises(pb.PumasError) as e:
        physics.property("range", 2, 0, 1.)  # STRAGGLED is not a table scheme
    assert e.value.rc == 7
    with pytest.raises(pb.PumasError) as e:
        physics.property("range", MODE_CSDA, 17, 1.)
    assert e.value.rc == 7


def test_load_errors(tmp_path):
    bad = tmp_path / "bad.pumas"
    bad.write_bytes(b"\x0c\x00\x00\x00" + b"\x00" * 64)  # tag 12
    with pytest.raises(pb.PumasError) as e:
        pb.Physics(str(bad))
    assert e.value.rc == 11  # FORMAT_ERROR, pumas.c:2077
    data = open(harness.dump_path(), "rb").read()
    cut = tmp_path / "cut.pumas"
    cut.write_bytes(data[:100000])
    with pytest.raises(pb.PumasError) as e:
        pb.Physics(str(cut))
    assert e.value.rc == 10  # IO_ERROR: truncated


def test_no_cpu_fallback(physics):
    """Without a B200 the transport entry point fails loudly."""
    if pb.device_available():
        pytest.skip("a GPU is present")
    ctx = pb.Context(physics)
    ctx.geometry(pb.uniform_geometry(0))
    with pytest.raises(pb.PumasError) as e:
        ctx.transport(pb.make_states(3, 1.))
    assert e.value.rc == 9  # PUMAS_RETURN_INTERNAL_ERROR


def test_error_catch_and_raise(physics):
    """pumas_error_catch / pumas_error_raise, pumas.c:2306-2327."""
    lib = pb.library()
    lib.pumas_error_catch(1)
    i = ctypes.c_int()
    rc = lib.pumas_physics_material_index(
        physics._handle, b"Nope", ctypes.byref(i))
    assert rc == 22
    rc2 = lib.pumas_physics_material_index(
        physics._handle, b"StandardRock", ctypes.byref(i))
    assert rc2 == 0
    assert lib.pumas_error_raise() == 22  # first caught error is re-raised
    assert lib.pumas_error_raise() == 18  # RAISE_ERROR: nothing was caught


def test_memory_hooks(physics):
    """pumas_memory_allocator / _deallocator (pumas.c:1379-1398): the objects
    handed to the user go through the hooks; NULL restores malloc / free."""
    import ctypes
    lib = pb.library()
    libc = ctypes.CDLL(None)
    libc.malloc.restype = ctypes.c_void_p
    libc.malloc.argtypes = [ctypes.c_size_t]
    libc.free.argtypes = [ctypes.c_void_p]
    calls = {"malloc": [], "free": []}

    @ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_size_t)
    def allocate(size):
        p = libc.malloc(size)
        calls["malloc"].append((p, size))
        return p

    @ctypes.CFUNCTYPE(None, ctypes.c_void_p)
    def deallocate(p):
        calls["free"].append(p)
        libc.free(p)

    lib.pumas_memory_allocator.argtypes = [ctypes.c_void_p]
    lib.pumas_memory_deallocator.argtypes = [ctypes.c_void_p]
    lib.pumas_memory_reallocator.argtypes = [ctypes.c_void_p]
    lib.pumas_memory_allocator(ctypes.cast(allocate, ctypes.c_void_p))
    lib.pumas_memory_deallocator(ctypes.cast(deallocate, ctypes.c_void_p))
    try:
        ctx = ctypes.POINTER(pb.ContextStruct)()
        assert lib.pumas_context_create(
            ctypes.byref(ctx), physics._handle, 64) == 0
        assert len(calls["malloc"]) == 1 and calls["malloc"][0][1] > 64
        base, size = calls["malloc"][0]
        user = ctypes.cast(ctx.contents.user_data, ctypes.c_void_p).value
        assert base < user and user + 64 <= base + size
        assert ctx.contents.accuracy == 1e-2  # zeroed, then the defaults
        lib.pumas_context_destroy(ctypes.byref(ctx))
        assert calls["free"] == [base] and not ctx

        other = pb.Physics(harness.dump_path())
        assert len(calls["malloc"]) == 2
        handle = other._handle.value
        assert handle == calls["malloc"][1][0]
        del other
        assert calls["free"][-1] == handle
    finally:
        lib.pumas_memory_allocator(None)
        lib.pumas_memory_deallocator(None)
        lib.pumas_memory_reallocator(None)
    n = len(calls["malloc"])
    ctx = pb.Context(physics)
    del ctx
    assert len(calls["malloc"]) == n


def test_random_dump_and_load(physics, tmp_path):
    """pumas_context_random_dump / _load (pumas.c:2415-2492): the state of the
    counter based streams is (seed, id of the next first track)."""
    import ctypes
    lib = pb.library()
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    for f in (lib.pumas_context_random_dump, lib.pumas_context_random_load):
        f.argtypes = [ctypes.POINTER(pb.ContextStruct), ctypes.c_void_p]
    lib.pumas_context_random_seed_get.argtypes = [
        ctypes.POINTER(pb.ContextStruct), ctypes.POINTER(ctypes.c_ulong)]

    a = pb.Context(physics, seed=123456789)
    a.track_offset(1 << 40)
    path = str(tmp_path / "random.bin").encode()
    stream = libc.fopen(path, b"wb")
    assert lib.pumas_context_random_dump(a._ptr, stream) == 0
    libc.fclose(stream)
    assert len(open(path, "rb").read()) == 4 + 16

    b = pb.Context(physics, seed=1)
    stream = libc.fopen(path, b"rb")
    assert lib.pumas_context_random_load(b._ptr, stream) == 0
    libc.fclose(stream)
    seed = ctypes.c_ulong()
    assert lib.pumas_context_random_seed_get(b._ptr, ctypes.byref(seed)) == 0
    assert seed.value == 123456789
    # and the offset: a second dump is byte identical
    other = str(tmp_path / "other.bin").encode()
    stream = libc.fopen(other, b"wb")
    assert lib.pumas_context_random_dump(b._ptr, stream) == 0
    libc.fclose(stream)
    assert open(other, "rb").read() == open(path, "rb").read()

    assert lib.pumas_context_random_dump(a._ptr, None) == 17  # PATH_ERROR
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"\x01\x00\x00\x00" + b"\x00" * 16)
    stream = libc.fopen(str(bad).encode(), b"rb")
    assert lib.pumas_context_random_load(b._ptr, stream) == 11  # FORMAT_ERROR
    libc.fclose(stream)
    bad.write_bytes(b"\x00\xb2\x00\x00" + b"\x00" * 4)
    stream = libc.fopen(str(bad).encode(), b"rb")
    assert lib.pumas_context_random_load(b._ptr, stream) == 10  # IO_ERROR
    libc.fclose(stream)


def test_composite_properties(physics, oracle):
    """pumas_physics_composite_properties (pumas.c:3446): WetRock of
    examples/data/materials.xml, against the reference."""
    import ctypes
    ours, ref = pb.library(), oracle.lib
    for lib in (ours, ref):
        lib.pumas_physics_composite_properties.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]
        lib.pumas_physics_composite_length.argtypes = [ctypes.c_void_p]
        lib.pumas_physics_material_length.argtypes = [ctypes.c_void_p]
    n = ours.pumas_physics_material_length(physics._handle)
    nc = ours.pumas_physics_composite_length(physics._handle)
    assert (n, nc) == (ref.pumas_physics_material_length(oracle.physics),
                       ref.pumas_physics_composite_length(oracle.physics))
    assert nc == 1
    for material in range(n - nc, n):
        got = []
        for lib, handle in ((ours, physics._handle), (ref, oracle.physics)):
            length = ctypes.c_int()
            assert lib.pumas_physics_composite_properties(
                handle, material, ctypes.byref(length), None, None) == 0
            components = (ctypes.c_int * length.value)()
            fractions = (ctypes.c_double * length.value)()
            assert lib.pumas_physics_composite_properties(
                handle, material, None, components, fractions) == 0
            got.append((length.value, list(components), list(fractions)))
        assert got[0] == got[1] and got[0][0] == 2
        assert abs(sum(got[0][2]) - 1.) < 1e-12
    # a base material is not a composite (pumas.c:3456-3459)
    assert ours.pumas_physics_composite_properties(
        physics._handle, 0, None, None, None) == 7
