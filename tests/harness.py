"""Shared test harness: ctypes bindings of the CPU oracle (oracle/_ref, the
reference itself), of the CPU debugging build of the core (oracle/_port) and the
workload definitions (BASELINE.json configs).

Only tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke() import this module.
"""
import ctypes
import gzip
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

import sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pumas_b200 import (  # noqa: E402,F401
    DENSITY_CONSTANT, DENSITY_EXPONENTIAL, EVENT_LIMIT_DISTANCE,
    EVENT_LIMIT_ENERGY, EVENT_LIMIT_GRAMMAGE, EVENT_LIMIT_TIME, EVENT_MEDIUM,
    EVENT_NONE, EVENT_VERTEX_COULOMB, EVENT_VERTEX_DECAY, EVENT_VERTEX_DEL,
    EVENT_WEIGHT, GEOMETRY_LAYERS, GEOMETRY_UNIFORM, MAX_LAYERS, MODE_BACKWARD,
    MODE_CSDA, MODE_DISABLED, MODE_FORWARD, MODE_MIXED, MODE_RANDOMISED,
    MODE_STRAGGLED, MODE_WEIGHTED, STATE_DTYPE, STEP_CHECK, STEP_RAW, Geometry,
    Layer, layered_geometry, make_states, uniform_geometry)
from pumas_b200 import ContextStruct as Context  # noqa: E402

STATE_FIELDS = ("charge", "energy", "distance", "grammage", "time", "weight",
                "position", "direction")


def make_settings(energy_loss=MODE_STRAGGLED, scattering=MODE_MIXED,
                  direction=MODE_FORWARD, decay=MODE_WEIGHTED, event=0,
                  limit_energy=0., limit_distance=0., limit_grammage=0.,
                  limit_time=0., accuracy=1e-2):
    c = Context()
    c.mode.energy_loss = energy_loss
    c.mode.scattering = scattering
    c.mode.direction = direction
    c.mode.decay = decay
    c.event = event
    c.limit.energy = limit_energy
    c.limit.distance = limit_distance
    c.limit.grammage = limit_grammage
    c.limit.time = limit_time
    c.accuracy = accuracy
    return c


# ---- fixtures ---------------------------------------------------------------
_dump_cache = {}


def dump_path(name="materials.muon"):
    """Path of the decompressed physics dump `name` (tests/golden/*.pumas.gz)."""
    if name not in _dump_cache:
        src = os.path.join(GOLDEN, name + ".pumas.gz")
        d = os.path.join(tempfile.gettempdir(), "pumas_b200_dumps_%d" % os.getuid())
        os.makedirs(d, exist_ok=True)
        dst = os.path.join(d, name + ".pumas")
        with gzip.open(src, "rb") as f:
            data = f.read()
        if not (os.path.exists(dst) and os.path.getsize(dst) == len(data)):
            tmp = dst + ".%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(data)
            os.replace(tmp, dst)
        _dump_cache[name] = dst
    return _dump_cache[name]


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctype))


# ---- the reference (oracle/_ref) -------------------------------------------
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "libpumas_oracle.so")
PORT_SO = os.path.join(ROOT, "oracle", "_port", "libpumas_port.so")
PORT_DD_SO = os.path.join(ROOT, "oracle", "_port", "libpumas_port_dd.so")


def have_oracle():
    return os.path.exists(ORACLE_SO)


def have_port():
    return os.path.exists(PORT_SO)


class Oracle:
    """The reference implementation (src/pumas.c compiled unmodified)."""

    def __init__(self, dump="materials.muon"):
        self.lib = lib = ctypes.CDLL(ORACLE_SO)
        lib.oracle_physics_load.argtypes = [
            ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]
        lib.oracle_transport_n.argtypes = [
            ctypes.c_void_p, ctypes.POINTER(Context), ctypes.POINTER(Geometry),
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_long, ctypes.c_void_p,
            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
            ctypes.POINTER(ctypes.c_uint), ctypes.c_int,
            ctypes.POINTER(ctypes.c_double)]
        lib.oracle_table.restype = ctypes.POINTER(ctypes.c_double)
        lib.oracle_table.argtypes = [
            ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long)]
        for name in ("oracle_dcs_evaluate", "oracle_dcs_model"):
            f = getattr(lib, name)
            f.restype = ctypes.c_double
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                          ctypes.c_double, ctypes.c_double]
        lib.oracle_coulomb_ehs_length.restype = ctypes.c_double
        lib.oracle_coulomb_ehs_length.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        lib.oracle_table_get_msc.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        lib.oracle_coulomb_cs_hard.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
            ctypes.POINTER(ctypes.c_double)]
        lib.pumas_physics_material_index.argtypes = [
            ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        self.physics = ctypes.c_void_p()
        rc = lib.oracle_physics_load(
            ctypes.byref(self.physics), dump_path(dump).encode())
        if rc != 0:
            raise RuntimeError("oracle_physics_load failed (rc=%d)" % rc)

    def material(self, name):
        i = ctypes.c_int(-1)
        rc = self.lib.pumas_physics_material_index(
            self.physics, name.encode(), ctypes.byref(i))
        if rc != 0:
            raise KeyError(name)
        return i.value

    def property(self, name, *args):
        """Call pumas_physics_property_<name>(physics, *args, &value)."""
        f = getattr(self.lib, "pumas_physics_property_" + name)
        value = ctypes.c_double()
        cargs = [ctypes.c_double(a) if isinstance(a, float) else ctypes.c_int(a)
                 for a in args]
        rc = f(self.physics, *cargs, ctypes.byref(value))
        return rc, value.value

    def table(self, name):
        n = ctypes.c_long()
        p = self.lib.oracle_table(self.physics, name.encode(), ctypes.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def transport_recorded(self, settings, geometry, states, period, seed=0,
                           offset=0, max_frames=1 << 22):
        """The reference with a pumas_recorder attached. Returns (states,
        events, frame_states, frame_media, frame_events, counts, rc)."""
        from pumas_b200 import STATE_DTYPE
        s = np.ascontiguousarray(states.copy())
        n = len(s)
        events = np.zeros(n, dtype=np.int32)
        fs = np.zeros(max_frames, dtype=STATE_DTYPE)
        fm = np.zeros(max_frames, dtype=np.int32)
        fe = np.zeros(max_frames, dtype=np.int32)
        counts = np.zeros(n, dtype=np.int32)
        f = self.lib.oracle_transport_recorded
        f.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
            ctypes.c_uint64, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p]
        rc = f(self.physics, ctypes.byref(settings), ctypes.byref(geometry),
               seed, offset, n, s.ctypes.data, events.ctypes.data, period,
               max_frames, fs.ctypes.data, fm.ctypes.data, fe.ctypes.data,
               counts.ctypes.data)
        total = int(counts.sum())
        assert total <= max_frames, "more frames than max_frames"
        return s, events, fs[:total], fm[:total], fe[:total], counts, rc

    def transport(self, settings, geometry, states, seed=0, offset=0,
                  threads=1):
        """Returns (states, events, media, n_random, rc, seconds)."""
        s = np.ascontiguousarray(states.copy())
        n = len(s)
        events = np.zeros(n, dtype=np.int32)
        media = np.zeros((n, 2), dtype=np.int32)
        nrand = np.zeros(n, dtype=np.uint32)
        seconds = ctypes.c_double()
        rc = self.lib.oracle_transport_n(
            self.physics, ctypes.byref(settings), ctypes.byref(geometry),
            seed, offset, n, s.ctypes.data, _ptr(events, ctypes.c_int),
            _ptr(media, ctypes.c_int), _ptr(nrand, ctypes.c_uint), threads,
            ctypes.byref(seconds))
        return s, events, media, nrand, rc, seconds.value


class Port:
    """CPU build of the single-source core (oracle/_port)."""

    def __init__(self, dump="materials.muon", dd=False):
        self.lib = lib = ctypes.CDLL(PORT_DD_SO if dd else PORT_SO)
        lib.port_physics_load.restype = ctypes.c_void_p
        lib.port_physics_load.argtypes = [ctypes.c_char_p]
        lib.port_transport_n.argtypes = [
            ctypes.c_void_p, ctypes.POINTER(Context), ctypes.POINTER(Geometry),
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_long, ctypes.c_void_p,
            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
            ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint),
            ctypes.c_int]
        lib.port_table.restype = ctypes.POINTER(ctypes.c_double)
        lib.port_table.argtypes = [
            ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long)]
        for name in ("port_dcs_evaluate", "port_dcs_model"):
            f = getattr(lib, name)
            f.restype = ctypes.c_double
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                          ctypes.c_double, ctypes.c_double]
        lib.port_coulomb_ehs_length.restype = ctypes.c_double
        lib.port_coulomb_ehs_length.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        lib.port_table_get_msc.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        lib.port_coulomb_cs_hard.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
            ctypes.POINTER(ctypes.c_double)]
        self.physics = lib.port_physics_load(dump_path(dump).encode())
        if not self.physics:
            raise RuntimeError("port_physics_load failed")

    def table(self, name):
        n = ctypes.c_long()
        p = self.lib.port_table(self.physics, name.encode(), ctypes.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def transport(self, settings, geometry, states, seed=0, offset=0,
                  threads=1):
        """Returns (states, events, media, n_random, n_steps, rc)."""
        s = np.ascontiguousarray(states.copy())
        n = len(s)
        events = np.zeros(n, dtype=np.int32)
        media = np.zeros((n, 2), dtype=np.int32)
        nrand = np.zeros(n, dtype=np.uint32)
        nsteps = np.zeros(n, dtype=np.uint32)
        rc = self.lib.port_transport_n(
            self.physics, ctypes.byref(settings), ctypes.byref(geometry),
            seed, offset, n, s.ctypes.data, _ptr(events, ctypes.c_int),
            _ptr(media, ctypes.c_int), _ptr(nrand, ctypes.c_uint),
            _ptr(nsteps, ctypes.c_uint), threads)
        return s, events, media, nrand, nsteps, rc


# ---- comparison ------------------------------------------------------------
def compare_states(a, b, events_a=None, events_b=None, nrand_a=None,
                   nrand_b=None, rtol=1e-9):
    """Per-track comparison of final states.

    Returns a dict with: n, n_match (tracks whose every field agrees within
    rtol relative, scaled by max(|a|, |b|, tiny)), n_diverged (tracks whose
    number of consumed variates or final event differ: the event sequence
    diverged), max_rel (largest relative difference over non diverged tracks),
    n_bitexact.
    """
    n = len(a)
    worst = np.zeros(n)
    bitexact = np.ones(n, dtype=bool)
    for f in STATE_FIELDS:
        x = a[f].reshape(n, -1)
        y = b[f].reshape(n, -1)
        scale = np.maximum(np.maximum(np.abs(x), np.abs(y)), 1e-300)
        with np.errstate(invalid="ignore"):
            rel = np.abs(x - y) / scale
        rel = np.where((x == y) | (np.isnan(x) & np.isnan(y)), 0., rel)
        # Components that are tiny w.r.t. the vector norm (transverse position
        # of a nearly straight track) are compared on the vector scale.
        if x.shape[1] == 3:
            norm = np.maximum(np.linalg.norm(x, axis=1), 1e-300)[:, None]
            with np.errstate(invalid="ignore"):
                rel = np.minimum(rel, np.abs(x - y) / norm)
            # (NaN on both sides is agreement: the reference itself returns
            # NaN for a few tracks, e.g. pumas.c:6781 under a time limit)
            rel = np.where((x == y) | (np.isnan(x) & np.isnan(y)), 0., rel)
        worst = np.maximum(worst, np.nan_to_num(rel, nan=np.inf).max(axis=1))
        bitexact &= ((x == y) | (np.isnan(x) & np.isnan(y))).all(axis=1)
    bitexact &= a["decayed"] == b["decayed"]
    diverged = np.zeros(n, dtype=bool)
    if events_a is not None:
        diverged |= events_a != events_b
    if nrand_a is not None:
        diverged |= nrand_a != nrand_b
    diverged |= a["decayed"] != b["decayed"]
    ok = (worst <= rtol) & ~diverged
    sync = ~diverged
    return {
        "n": n,
        "n_match": int(ok.sum()),
        "n_diverged": int(diverged.sum()),
        "n_bitexact": int((bitexact & sync).sum()),
        "max_rel": float(worst[sync].max()) if sync.any() else 0.,
        "worst": worst,
        "diverged": diverged,
    }
