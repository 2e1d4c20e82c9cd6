"""Recorded outputs of the reference (tests/golden/tracks.npz, made by
tests/golden/make_golden.py) against the CPU build of the core (bit exact,
no GPU) and against the GPU through the C ABI (1e-9)."""
import os

import numpy as np
import pytest

import harness
from workloads import WORKLOADS

N, SEED, OFFSET = 48, 21, 4096
GOLDEN = np.load(os.path.join(harness.GOLDEN, "tracks.npz"))


def golden(name):
    return (GOLDEN[name + ".states"], GOLDEN[name + ".events"],
            GOLDEN[name + ".media"], GOLDEN[name + ".nrand"])


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_port_reproduces_golden(name):
    if not harness.have_port():
        pytest.skip("oracle/_port not built")
    dump, builder, kwargs = WORKLOADS[name]
    import pumas_b200 as pb
    physics = pb.Physics(harness.dump_path(dump))
    port = harness.Port(dump)
    w = builder(physics.material_index, N, **kwargs)
    s, e, m, r, steps, rc = port.transport(
        w["settings"], w["geometry"], w["states"], seed=SEED, offset=OFFSET)
    gs, ge, gm, gr = golden(name)
    c = harness.compare_states(gs, s, ge, e, gr, r)
    assert c["n_bitexact"] == N and c["n_diverged"] == 0
    assert np.array_equal(gm, m)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_gpu_reproduces_golden(name):
    import pumas_b200 as pb
    dump, builder, kwargs = WORKLOADS[name]
    physics = pb.Physics(harness.dump_path(dump))
    w = builder(physics.material_index, N, **kwargs)
    ctx = pb.Context(physics, seed=SEED)
    ctx.configure(w["settings"])
    ctx.geometry(w["geometry"])
    ctx.track_offset(OFFSET)
    s = w["states"].copy()
    e, m = ctx.transport(s)
    r, _ = ctx.last_counters(N)
    gs, ge, gm, gr = golden(name)
    c = harness.compare_states(gs, s, ge, e, gr, r, rtol=1e-9)
    assert c["n_match"] >= N - 1, (c["n_match"], c["n_diverged"], c["max_rel"])
