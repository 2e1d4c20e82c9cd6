"""The track-resident kernel against the monolithic one (one thread per track,
whole loop) on batches large enough for the scheduler's steady state: more
tracks than lanes, so that the work cursor, the warp pools, the couriers and
the DEL CTAs (resident.cuh, PB_RS_OFFLOAD) all run for a while before the
batch ends. Results must not depend on the schedule: same per-track Philox
streams, bit-identical states, events, media and numbers of variates.

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


@pytest.mark.parametrize("name,n", CASES)
def test_resident_equals_monolithic(name, n):
    assert pb.device_available(), "no B200 visible: the GPU tests need one"
    dump, builder, kwargs = WORKLOADS[name]
    physics = pb.Physics(harness.dump_path(dump))
    w = builder(physics.material_index, n, **kwargs)
    a = _run(physics, w, 0)
    b = _run(physics, w, 1)
    for f in harness.STATE_FIELDS:
        x, y = a[0][f], b[0][f]
        same = (x == y) | (np.isnan(x) & np.isnan(y))
        assert same.all(), (name, f, int((~same).sum()))
    assert np.array_equal(a[1], b[1]), "events"
    assert np.array_equal(a[2], b[2]), "media"
    assert np.array_equal(a[3], b[3]), "variates drawn"
    assert np.array_equal(a[4], b[4]), "steps"
    # and twice the same batch gives twice the same answer (the schedule of
    # the first run is not the schedule of the second)
    c = _run(physics, w, 0)
    for f in harness.STATE_FIELDS:
        x, y = a[0][f], c[0][f]
        assert ((x == y) | (np.isnan(x) & np.isnan(y))).all(), (name, f)
    assert np.array_equal(a[3], c[3])
