"""The recorder (pumas.h:432-458; hooks pumas.c:3191, 4798, 4946, 5230, 5326,
5368, 5385, 5466, 5509): frames of the restated engine against the frames the
reference itself records, track by track.

CPU: the engine's hooks with a host recorder (oracle/_port) -- bit exact.
GPU: the device recorder through the C ABI (pumas_recorder_create attached to
context->recorder, pumas_context_transport_n) -- same frames (number, events,
media), states within 1e-9."""
import ctypes

import numpy as np
import pytest

import harness
from workloads import WORKLOADS

CASES = [("geometry_detailed", 40, 1), ("geometry_detailed", 40, 7),
         ("geometry_detailed", 40, 0), ("headline", 6, 50),
         ("straight_mixed", 30, 3), ("csda_water", 50, 1),
         ("layers_check", 40, 2), ("mixed_loss_scattering", 30, 5)]


def _reference(name, n, period):
    if not harness.have_oracle():
        pytest.skip("oracle/_ref not built")
    dump, builder, kwargs = WORKLOADS[name]
    oracle = harness.Oracle(dump)
    w = builder(oracle.material, n, **kwargs)
    return w, oracle.transport_recorded(
        w["settings"], w["geometry"], w["states"], period, seed=9, offset=77)


def _same_frames(ref, got, exact):
    _, _, fs, fm, fe, counts, _ = ref
    gs, gm, ge, gcounts = got
    assert np.array_equal(counts, gcounts), "frames per track"
    assert np.array_equal(fe, ge), "events of the frames"
    assert np.array_equal(fm, gm), "media of the frames"
    c = harness.compare_states(fs, gs, rtol=1e-9)
    if exact:
        assert c["n_bitexact"] == len(fs)
    else:
        assert c["n_match"] >= len(fs) - max(1, len(fs) // 2000), c["max_rel"]


@pytest.mark.parametrize("name,n,period", CASES)
def test_port_frames_bit_exact(name, n, period):
    if not harness.have_port():
        pytest.skip("oracle/_port not built")
    w, ref = _reference(name, n, period)
    import pumas_b200 as pb
    dump = WORKLOADS[name][0]
    port = harness.Port(dump)
    s = np.ascontiguousarray(w["states"].copy())
    events = np.zeros(n, dtype=np.int32)
    total = int(ref[5].sum())
    fs = np.zeros(total + 16, dtype=pb.STATE_DTYPE)
    fm = np.zeros(total + 16, dtype=np.int32)
    fe = np.zeros(total + 16, dtype=np.int32)
    counts = np.zeros(n, dtype=np.int32)
    f = port.lib.port_transport_recorded
    f.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
        ctypes.c_uint64, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]
    rc = f(port.physics, ctypes.byref(w["settings"]),
           ctypes.byref(w["geometry"]), 9, 77, n, s.ctypes.data,
           events.ctypes.data, period, total + 16, fs.ctypes.data,
           fm.ctypes.data, fe.ctypes.data, counts.ctypes.data)
    assert rc == 0
    assert int(counts.sum()) == total
    _same_frames(ref, (fs[:total], fm[:total], fe[:total], counts), True)
    # the recorder changes nothing to the transport itself
    c = harness.compare_states(ref[0], s, ref[1], events)
    assert c["n_bitexact"] == n


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,period", CASES)
def test_gpu_frames(name, n, period):
    import pumas_b200 as pb
    w, ref = _reference(name, n, period)
    physics = pb.Physics(harness.dump_path(WORKLOADS[name][0]))
    ctx = pb.Context(physics, seed=9)
    ctx.configure(w["settings"])
    ctx.geometry(w["geometry"])
    ctx.track_offset(77)
    ctx.record(period)
    s = w["states"].copy()
    events, _ = ctx.transport(s)
    gs, gm, ge = ctx.frames()
    # frames are appended in track order: a track starts at a START frame
    starts = np.flatnonzero(ge & 4096)
    counts = np.diff(np.append(starts, len(ge))).astype(np.int32)
    assert len(starts) == n
    _same_frames(ref, (gs, gm, ge, counts), False)
    c = harness.compare_states(ref[0], s, ref[1], events, rtol=1e-9)
    assert c["n_match"] == n
    # without the recorder: the product kernels, same final states
    ctx.record(None)
    ctx.track_offset(77)
    s2 = w["states"].copy()
    e2, _ = ctx.transport(s2)
    c = harness.compare_states(s, s2, events, e2, rtol=1e-9)
    assert c["n_match"] == n


def test_recorder_objects():
    """pumas_recorder_create / clear / destroy and the defaults of
    pumas.c:2816-2820."""
    import pumas_b200 as pb
    lib = pb.library()
    r = ctypes.c_void_p()
    assert lib.pumas_recorder_create(ctypes.byref(r), 16) == 0
    assert ctypes.c_void_p.from_address(r.value).value is None  # first
    assert ctypes.c_int.from_address(r.value + 8).value == 0    # length
    assert ctypes.c_int.from_address(r.value + 12).value == 1   # period
    assert ctypes.c_void_p.from_address(r.value + 24).value     # user_data
    lib.pumas_recorder_clear(r)
    lib.pumas_recorder_destroy(ctypes.byref(r))
    assert r.value is None
