"""examples/straight_b200.c: a C user program written against include/*.h
links with libpumas_b200.so (CPU check), runs on the GPU and prints the same
estimate as the Python binding on the same batch (GPU check)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import harness
import pumas_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pumas_b200")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    if not os.path.exists(os.path.join(LIBDIR, "libpumas_b200.so")):
        pytest.fail("pumas_b200/libpumas_b200.so is not built")
    exe = str(tmp_path / "straight_b200")
    subprocess.run(
        ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror",
         os.path.join(ROOT, "examples", "straight_b200.c"),
         "-I" + os.path.join(ROOT, "include"), "-L" + LIBDIR, "-lpumas_b200",
         "-lm", "-Wl,-rpath," + LIBDIR, "-o", exe], check=True)
    return exe


def test_example_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    # without arguments: usage on stderr, no library call
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "usage" in r.stderr


@pytest.mark.gpu
def test_example_matches_python_binding(tmp_path):
    exe = _build(tmp_path)
    dump = harness.dump_path("materials.muon")
    n, depth, kinetic = 20000, 1000., 1.
    r = subprocess.run([exe, dump, str(depth), str(kinetic), str(kinetic),
                        str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"Flux : (\S+) \\pm (\S+)", r.stdout)
    assert m, r.stdout
    flux, sigma = float(m.group(1)), float(m.group(2))

    physics = pb.Physics(dump)
    ctx = pb.Context(physics, seed=0)
    ctx.api.mode.energy_loss = pb.MODE_STRAGGLED
    ctx.api.mode.direction = pb.MODE_BACKWARD
    ctx.api.mode.scattering = pb.MODE_DISABLED
    ctx.api.limit.distance = depth
    ctx.api.event = pb.EVENT_LIMIT_DISTANCE
    ctx.geometry(pb.uniform_geometry(physics.material_index("StandardRock")))
    states = pb.make_states(n, kinetic)
    states["weight"] = 2.
    states["direction"][:, 2] = -1.
    events, _ = ctx.transport(states)
    ok = (events & pb.EVENT_LIMIT_DISTANCE) != 0
    x = np.where(ok, states["weight"] * (states["energy"] + 0.10566) ** -2.7,
                 0.)
    # the charge (drawn by the C program) does not enter without a field
    assert flux > 0. and sigma > 0.
    assert abs(flux - x.mean()) <= 1e-5 * flux  # 6 printed digits
