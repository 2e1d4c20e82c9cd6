"""`__device__` medium / locals callbacks through the C ABI (SURVEY.md section
8f.3; pumas.h:509-511, 348-349):
  - the library's own client (the planar media behind the callback interface)
    gives bit for bit the results of the descriptor path;
  - user functions (examples/user_media.cu: spherical shells with radial
    density gradients), device-linked by `python -m pumas_b200.build --user`,
    agree with the unmodified reference driving the same geometry through host
    callbacks (oracle/_ref)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import harness
import pumas_b200 as pb
from workloads import WORKLOADS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
USER_SO = os.path.join(ROOT, "examples", "libpumas_b200_user.so")


def _user_library():
    if not os.path.exists(USER_SO):
        from pumas_b200 import build
        build.build_user(
            [os.path.join(ROOT, "examples", "user_media.cu")], USER_SO)
    return USER_SO


def test_user_library_exports():
    """The device-linked user build holds the whole ABI and the user's entry."""
    so = _user_library()
    out = subprocess.run(["nm", "-D", "--defined-only", so],
                         capture_output=True, text=True, check=True).stdout
    names = {l.split()[-1] for l in out.splitlines() if l.strip()}
    for symbol in ("user_media_callbacks", "pumas_b200_context_callbacks_set",
                   "pumas_b200_planar_callbacks", "pumas_context_transport_n",
                   "pumas_context_transport", "pumas_physics_load"):
        assert symbol in names, symbol


def test_callbacks_validation():
    """Argument checks need no device."""
    physics = pb.Physics(harness.dump_path("materials.muon"))
    ctx = pb.Context(physics)
    cb = pb.Callbacks()
    with pytest.raises(pb.PumasError) as e:  # no medium function
        ctx.callbacks(cb)
    assert e.value.rc == 12
    cb.medium = 1
    cb.n_media = 0
    with pytest.raises(pb.PumasError) as e:
        ctx.callbacks(cb)
    assert e.value.rc == 12
    cb.n_media = 1
    cb.material[0] = 99
    with pytest.raises(pb.PumasError) as e:
        ctx.callbacks(cb)
    assert e.value.rc == 7
    cb.material[0] = 0
    cb.has_locals[0] = 1  # locals flagged, no function
    with pytest.raises(pb.PumasError) as e:
        ctx.callbacks(cb)
    assert e.value.rc == 12
    ctx.callbacks(None)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["geometry_detailed", "geometry_fast",
                                  "layers_check", "headline",
                                  "limits_all_csda", "decay_randomised"])
def test_planar_callbacks_equal_descriptor(name):
    assert pb.device_available()
    dump, builder, kwargs = WORKLOADS[name]
    physics = pb.Physics(harness.dump_path(dump))
    n = 20000
    w = builder(physics.material_index, n, **kwargs)

    def run(callbacks):
        ctx = pb.Context(physics, seed=5)
        ctx.configure(w["settings"])
        if callbacks is None:
            ctx.geometry(w["geometry"])
            # the general resident kernel: same code as the callback kernel
            # but for the media (the shared memory kernel of the headline has
            # its own elementary functions, <= 1 ulp apart)
            ctx.kernel(2)
        else:
            ctx.callbacks(callbacks)
        ctx.track_offset(77)
        s = w["states"].copy()
        events, media = ctx.transport(s)
        return s, events, media, ctx.last_counters(n)[0]

    s0, e0, m0, r0 = run(None)
    cb = pb.planar_callbacks(w["geometry"])
    s1, e1, m1, r1 = run(cb)
    pb.library().pumas_b200_planar_callbacks_free(ctypes.byref(cb))
    assert np.array_equal(e0, e1)
    assert np.array_equal(m0, m1)
    assert np.array_equal(r0, r1)
    # bit for bit (NaN included: the reference itself returns NaN for a few
    # tracks under a time limit, pumas.c:6781)
    for f in harness.STATE_FIELDS + ("decayed",):
        assert np.ascontiguousarray(s0[f]).tobytes() == \
            np.ascontiguousarray(s1[f]).tobytes(), f


@pytest.mark.gpu
def test_user_media_vs_reference():
    if not harness.have_oracle():
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, PUMAS_B200_LIBRARY=_user_library())
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "tests", "_user_media_run.py"),
         "media", "20000"], capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    print(r)
    for name, c in r.items():
        assert c["rc"] == 0, (name, c)
        assert c["media_equal"], (name, c)
        # same criterion as tests/test_gpu_parity.py: every field of every
        # track within 1e-9 -- but the direction of the forward tracks that
        # come to rest inside the density gradient: their last steps scatter
        # by large angles scaled by step lengths that differ in the last bits
        # (measured on B200: 770 of 20 000 such directions off by <= 3.4e-7,
        # all other fields <= 4e-10)
        assert c["diverged"] == 0, (name, c)
        assert c["off_but_rest_direction"] == 0, (name, c)
        assert c["rest_direction_max"] <= 1e-5, (name, c)
        assert c["media_changes"] + c["exits"] > 0, (name, c)


def test_user_dcs_host_side():
    """pumas_dcs_register / _get / pumas_physics_load / pumas_physics_dcs with
    a user model, as the reference behaves (pumas.c:15422-15500): no device
    needed. Runs in a subprocess: the registry is process wide."""
    code = r"""
import ctypes, sys
sys.path.insert(0, 'tests')
import harness, pumas_b200 as pb
lib = pb.library()
try:
    pb.Physics(harness.dump_path('rock.muon.usr'))
    raise SystemExit('loaded without the model')
except pb.PumasError as e:
    assert e.rc == 14, e.rc
assert lib.user_dcs_register_host() == 0
assert lib.user_dcs_register_host() == 14   # already registered
DCS = ctypes.CFUNCTYPE(ctypes.c_double, *([ctypes.c_double] * 5))
f = ctypes.c_void_p()
lib.pumas_dcs_get.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p]
assert lib.pumas_dcs_get(0, b'BHS', ctypes.byref(f)) == 0
v = DCS(f.value)(11., 22., 0.10566, 100., 10.)
assert 1e-34 < v < 1e-29, v
ph = pb.Physics(harness.dump_path('rock.muon.usr'))
name = ctypes.c_char_p(); g = ctypes.c_void_p()
lib.pumas_physics_dcs.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
assert lib.pumas_physics_dcs(ph._handle, 0, ctypes.byref(name), ctypes.byref(g)) == 0
assert name.value == b'BHS' and g.value == f.value
# the reference gives the same value for the same model
o = ctypes.CDLL(harness.ORACLE_SO)
assert o.oracle_register_user_dcs() == 0
h = ctypes.c_void_p()
o.pumas_dcs_get.argtypes = lib.pumas_dcs_get.argtypes
assert o.pumas_dcs_get(0, b'BHS', ctypes.byref(h)) == 0
assert DCS(h.value)(11., 22., 0.10566, 100., 10.) == v
print('ok')
"""
    if not harness.have_oracle():
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, PUMAS_B200_LIBRARY=_user_library())
    out = subprocess.run([sys.executable, "-c", code], capture_output=True,
                         text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("ok")


@pytest.mark.gpu
def test_user_dcs_vs_reference():
    """Transport with a user registered DCS model (device function called by
    the device-linked kernel) against the reference with the same model."""
    if not harness.have_oracle():
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, PUMAS_B200_LIBRARY=_user_library())
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "tests", "_user_media_run.py"),
         "dcs", "20000"], capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    print(r)
    assert r.pop("unregistered_rc") == 14
    for name, c in r.items():
        assert c["rc"] == 0, (name, c)
        assert c["diverged"] <= 2e-3 * c["n"], (name, c)
        assert c["off_in_sync"] <= 2e-3 * c["n"], (name, c)
        assert c["match"] >= 0.998 * c["n"], (name, c)


def _build_shells(tmp_path):
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    so = _user_library()
    exe = str(tmp_path / "shells_b200")
    subprocess.run(
        ["gcc", "-std=c99", "-D_DEFAULT_SOURCE", "-Wall", "-Wextra", "-Werror",
         os.path.join(ROOT, "examples", "shells_b200.c"),
         "-I" + os.path.join(ROOT, "include"),
         "-I" + os.path.join(ROOT, "examples"),
         "-L" + os.path.dirname(so), "-lpumas_b200_user", "-lm",
         "-Wl,-rpath," + os.path.dirname(so), "-o", exe], check=True)
    return exe


def test_c_program_on_user_media_links(tmp_path):
    """examples/shells_b200.c: a C user program against include/*.h and the
    user's device-linked library."""
    exe = _build_shells(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "usage" in r.stderr


@pytest.mark.gpu
def test_c_program_on_user_media_matches_python_binding(tmp_path):
    import re
    exe = _build_shells(tmp_path)
    n, kinetic = 10000, 150.
    r = subprocess.run([exe, harness.dump_path("materials.muon"), str(n),
                        str(kinetic)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    left = int(re.search(r"Left\s+: (\d+)", r.stdout).group(1))
    stopped = [int(x) for x in re.search(
        r"Stopped : (\d+) in rock, (\d+) in water, (\d+) in air",
        r.stdout).groups()]
    path = float(re.search(r"Path\s+: (\S+) m", r.stdout).group(1))
    assert left + sum(stopped) == n and left > 0 and sum(stopped) > 0

    code = r"""
import ctypes, sys
sys.path.insert(0, 'tests')
import numpy as np
import harness, pumas_b200 as pb
import _user_media_run as um
n, kinetic = %d, %r
physics = pb.Physics(harness.dump_path('materials.muon'))
g, materials = um.shells(physics.material_index)
lib = pb.library()
cb = pb.Callbacks()
lib.user_media_callbacks.argtypes = [ctypes.POINTER(um.UserMedia),
    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(pb.Callbacks)]
assert lib.user_media_callbacks(ctypes.byref(g), materials, ctypes.byref(cb)) == 0
ctx = pb.Context(physics, seed=1)
ctx.configure(harness.make_settings(pb.MODE_STRAGGLED, pb.MODE_MIXED))
ctx.callbacks(cb)
phi = 2. * np.pi * (np.arange(n) + 0.5) / n
d = np.stack([np.cos(phi), np.zeros(n), np.sin(phi)], axis=1)
s = pb.make_states(n, kinetic, direction=d, position=(10., 5., -20.))
ev, md = ctx.transport(s)
print(int((md[:, 1] < 0).sum()), '%%.6f' %% s['distance'].mean())
""" % (n, kinetic)
    env = dict(os.environ, PUMAS_B200_LIBRARY=_user_library())
    out = subprocess.run([sys.executable, "-c", code], capture_output=True,
                         text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    left_py, path_py = out.stdout.split()
    assert int(left_py) == left
    assert abs(float(path_py) - path) <= 2e-6
