"""The persistent kernels against the monolithic one (one thread per track,
whole loop) on batches large enough for the schedulers' steady state: more
tracks than lanes, so that the work cursor, the slots / warp pools, the service
and DEL CTAs (hot_kernel.cuh, resident.cuh) all run for a while before the
batch ends. Results must not depend on the schedule: same per-track Philox
streams, hence
  - the register resident kernel (kernel 3) == the monolithic one (kernel 1)
    bit for bit: same arithmetic, different schedule;
  - the shared memory resident kernel (kernel 0, straggled mode in a uniform
    medium) takes its logarithms and sines from pb_hot_math.h (<= 2 ulp from
    CUDA's): same event sequence track by track (events, numbers of variates
    and of steps) and states within 1e-9 but for a few tracks in 1e5; and it
    reproduces itself bit for bit from one run to the next.

(test_gpu_parity.py compares the CUDA path with the reference on small
batches, where every track is taken from the cursor at once.)
"""
import numpy as np
import pytest

import harness
import pumas_b200 as pb
from workloads import WORKLOADS

pytestmark = pytest.mark.gpu

CASES = [("headline", 150000), ("straight_mixed", 150000),
         ("rock_3km", 60000), ("straight", 200000), ("geometry_detailed", 30000)]


def _run(physics, w, kernel, seed=5, offset=12345):
    ctx = pb.Context(physics, seed=seed)
    ctx.configure(w["settings"])
    ctx.geometry(w["geometry"])
    ctx.track_offset(offset)
    ctx.kernel(kernel)
    states = w["states"].copy()
    events, media = ctx.transport(states)
    nrand, nsteps = ctx.last_counters(len(states))
    return states, events, media, nrand, nsteps


HOT = ("headline", "straight_mixed", "rock_3km", "straight")


def _same(a, b):
    for f in harness.STATE_FIELDS:
        x, y = a[0][f], b[0][f]
        same = (x == y) | (np.isnan(x) & np.isnan(y))
        assert same.all(), (f, int((~same).sum()))
    assert np.array_equal(a[1], b[1]), "events"
    assert np.array_equal(a[2], b[2]), "media"
    assert np.array_equal(a[3], b[3]), "variates drawn"
    assert np.array_equal(a[4], b[4]), "steps"


@pytest.mark.parametrize("name,n", CASES)
def test_resident_equals_monolithic(name, n):
    assert pb.device_available(), "no B200 visible: the GPU tests need one"
    dump, builder, kwargs = WORKLOADS[name]
    physics = pb.Physics(harness.dump_path(dump))
    w = builder(physics.material_index, n, **kwargs)
    mono = _run(physics, w, 1)
    _same(_run(physics, w, 3), mono)
    a = _run(physics, w, 0)
    if name in HOT:
        # same event sequence for (nearly) every track; states within 1e-9
        c = harness.compare_states(mono[0], a[0], mono[1], a[1], mono[3], a[3])
        print("\n%s: hot kernel vs monolithic: diverged=%d match=%d/%d "
              "max_rel=%.3g" % (name, c["n_diverged"], c["n_match"], n,
                                c["max_rel"]))
        assert c["n_diverged"] <= 2e-5 * n
        assert c["n"] - c["n_match"] <= 1e-4 * n
        sync = ~c["diverged"]
        assert np.array_equal(a[4][sync], mono[4][sync]), "steps"
        assert np.array_equal(a[2], mono[2]), "media"
    else:
        _same(a, mono)
    # and twice the same batch gives twice the same answer (the schedule of
    # the first run is not the schedule of the second)
    _same(_run(physics, w, 0), a)
