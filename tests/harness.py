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

# ---- enums (include/pumas_b200.h) -----------------------------------------
MODE_DISABLED, MODE_CSDA, MODE_MIXED, MODE_STRAGGLED = -1, 0, 1, 2
MODE_WEIGHTED, MODE_RANDOMISED = 0, 1
MODE_FORWARD, MODE_BACKWARD = 0, 1
EVENT_NONE = 0
EVENT_LIMIT_ENERGY, EVENT_LIMIT_DISTANCE = 1, 2
EVENT_LIMIT_GRAMMAGE, EVENT_LIMIT_TIME = 4, 8
EVENT_MEDIUM = 16
EVENT_VERTEX_DEL, EVENT_VERTEX_COULOMB, EVENT_VERTEX_DECAY = 480, 512, 1024
EVENT_WEIGHT = 2048
STEP_CHECK, STEP_RAW = 0, 1
GEOMETRY_UNIFORM, GEOMETRY_LAYERS = 0, 1
DENSITY_CONSTANT, DENSITY_EXPONENTIAL = 0, 1
MAX_LAYERS = 8

# struct pumas_state (pumas.h:289-309): 104 bytes.
STATE_DTYPE = np.dtype([
    ("charge", "f8"), ("energy", "f8"), ("distance", "f8"), ("grammage", "f8"),
    ("time", "f8"), ("weight", "f8"), ("position", "f8", 3),
    ("direction", "f8", 3), ("decayed", "i4"), ("_pad", "i4")])
assert STATE_DTYPE.itemsize == 104
STATE_FIELDS = ("charge", "energy", "distance", "grammage", "time", "weight",
                "position", "direction")


class ContextMode(ctypes.Structure):
    _fields_ = [("energy_loss", ctypes.c_int), ("decay", ctypes.c_int),
                ("direction", ctypes.c_int), ("scattering", ctypes.c_int)]


class ContextLimit(ctypes.Structure):
    _fields_ = [("energy", ctypes.c_double), ("distance", ctypes.c_double),
                ("grammage", ctypes.c_double), ("time", ctypes.c_double)]


class Context(ctypes.Structure):
    """struct pumas_context (pumas.h:636-672)."""
    _fields_ = [("medium", ctypes.c_void_p), ("random", ctypes.c_void_p),
                ("recorder", ctypes.c_void_p), ("user_data", ctypes.c_void_p),
                ("mode", ContextMode), ("event", ctypes.c_int),
                ("limit", ContextLimit), ("accuracy", ctypes.c_double)]


class Layer(ctypes.Structure):
    _fields_ = [("material", ctypes.c_int), ("has_locals", ctypes.c_int),
                ("density_model", ctypes.c_int), ("reserved", ctypes.c_int),
                ("density", ctypes.c_double), ("scale", ctypes.c_double),
                ("eps", ctypes.c_double), ("magnet", ctypes.c_double * 3)]


class Geometry(ctypes.Structure):
    """struct pumas_b200_geometry (include/pumas_b200_geometry.h)."""
    _fields_ = [("kind", ctypes.c_int), ("n_layers", ctypes.c_int),
                ("step_type", ctypes.c_int), ("reserved", ctypes.c_int),
                ("step_default", ctypes.c_double),
                ("step_epsilon", ctypes.c_double),
                ("z", ctypes.c_double * (MAX_LAYERS + 1)),
                ("layer", Layer * MAX_LAYERS)]


def uniform_geometry(material, density=None):
    """One infinite medium (examples/pumas/straight.c:61-74)."""
    g = Geometry()
    g.kind = GEOMETRY_UNIFORM
    g.n_layers = 1
    g.step_type = STEP_CHECK
    g.step_default = 1e3
    g.step_epsilon = 1e-7
    g.layer[0].material = material
    if density is not None:
        g.layer[0].has_locals = 1
        g.layer[0].density = density
    return g


def layered_geometry(z, layers, step_type=STEP_RAW):
    """Stack of horizontal layers (examples/pumas/geometry.c:104-163).
    layers: list of dicts(material, density=None, scale=None, eps=0.05,
    magnet=(0, 0, 0))."""
    assert len(z) == len(layers) + 1 <= MAX_LAYERS + 1
    g = Geometry()
    g.kind = GEOMETRY_LAYERS
    g.n_layers = len(layers)
    g.step_type = step_type
    g.step_default = 1e3
    g.step_epsilon = 1e-7
    for i, zi in enumerate(z):
        g.z[i] = zi
    for i, l in enumerate(layers):
        g.layer[i].material = l["material"]
        if l.get("density") is not None:
            g.layer[i].has_locals = 1
            g.layer[i].density = l["density"]
            if l.get("scale") is not None:
                g.layer[i].density_model = DENSITY_EXPONENTIAL
                g.layer[i].scale = l["scale"]
                g.layer[i].eps = l.get("eps", 0.05)
            for k, b in enumerate(l.get("magnet", (0., 0., 0.))):
                g.layer[i].magnet[k] = b
    return g


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


def make_states(n, energy, charge=-1., direction=(0., 0., 1.), weight=1.,
                position=(0., 0., 0.)):
    s = np.zeros(n, dtype=STATE_DTYPE)
    s["charge"] = charge
    s["energy"] = energy
    s["weight"] = weight
    s["direction"] = np.asarray(direction, dtype="f8")
    s["position"] = np.asarray(position, dtype="f8")
    return s


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
            rel = np.minimum(rel, np.abs(x - y) / norm)
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
