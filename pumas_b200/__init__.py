"""pumas_b200 -- thin ctypes binding of libpumas_b200.so.

The product is the C-ABI shared library (include/pumas_b200.h): hand-written
sm_100a CUDA kernels behind the `pumas.h` surface of the reference plus the
batched `pumas_context_transport_n`. This module only mirrors that C API for
Python callers (tests, bench.py); it contains no transport code and no CPU
fallback: if the library is missing, or no B200 is visible, calls fail loudly.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY = os.environ.get(
    "PUMAS_B200_LIBRARY", os.path.join(HERE, "libpumas_b200.so"))

# ---- enums (include/pumas_b200.h; pumas.h:90-236) ---------------------------
MODE_DISABLED, MODE_CSDA, MODE_MIXED, MODE_STRAGGLED = -1, 0, 1, 2
MODE_WEIGHTED, MODE_RANDOMISED = 0, 1
MODE_FORWARD, MODE_BACKWARD = 0, 1
EVENT_NONE = 0
EVENT_LIMIT_ENERGY, EVENT_LIMIT_DISTANCE = 1, 2
EVENT_LIMIT_GRAMMAGE, EVENT_LIMIT_TIME = 4, 8
EVENT_LIMIT = 15
EVENT_MEDIUM = 16
EVENT_VERTEX_BREMSSTRAHLUNG, EVENT_VERTEX_PAIR_CREATION = 32, 64
EVENT_VERTEX_PHOTONUCLEAR, EVENT_VERTEX_DELTA_RAY = 128, 256
EVENT_VERTEX_DEL, EVENT_VERTEX_COULOMB, EVENT_VERTEX_DECAY = 480, 512, 1024
EVENT_VERTEX = 2016
EVENT_WEIGHT = 2048
STEP_CHECK, STEP_RAW = 0, 1
GEOMETRY_UNIFORM, GEOMETRY_LAYERS = 0, 1
DENSITY_CONSTANT, DENSITY_EXPONENTIAL = 0, 1
MAX_LAYERS = 8
RETURN_SUCCESS = 0
RETURN_NAMES = (
    "SUCCESS", "ACCURACY_ERROR", "END_OF_FILE", "DECAY_ERROR", "DENSITY_ERROR",
    "DIRECTION_ERROR", "INCOMPLETE_FILE", "INDEX_ERROR", "PHYSICS_ERROR",
    "INTERNAL_ERROR", "IO_ERROR", "FORMAT_ERROR", "MEDIUM_ERROR",
    "MEMORY_ERROR", "MODEL_ERROR", "MISSING_LIMIT", "MISSING_RANDOM",
    "PATH_ERROR", "RAISE_ERROR", "TOO_LONG", "UNDEFINED_MDF",
    "UNKNOWN_ELEMENT", "UNKNOWN_MATERIAL", "UNKNOWN_PARTICLE", "VALUE_ERROR")

# struct pumas_state (pumas.h:289-309): 104 bytes.
STATE_DTYPE = np.dtype([
    ("charge", "f8"), ("energy", "f8"), ("distance", "f8"), ("grammage", "f8"),
    ("time", "f8"), ("weight", "f8"), ("position", "f8", 3),
    ("direction", "f8", 3), ("decayed", "i4"), ("_pad", "i4")])
assert STATE_DTYPE.itemsize == 104


class ContextMode(ctypes.Structure):
    _fields_ = [("energy_loss", ctypes.c_int), ("decay", ctypes.c_int),
                ("direction", ctypes.c_int), ("scattering", ctypes.c_int)]


class ContextLimit(ctypes.Structure):
    _fields_ = [("energy", ctypes.c_double), ("distance", ctypes.c_double),
                ("grammage", ctypes.c_double), ("time", ctypes.c_double)]


class ContextStruct(ctypes.Structure):
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


class Callbacks(ctypes.Structure):
    """struct pumas_b200_callbacks (include/pumas_b200_geometry.h): device
    function pointers, filled by the library that holds the functions."""
    _fields_ = [("medium", ctypes.c_void_p), ("locals", ctypes.c_void_p),
                ("user_data", ctypes.c_void_p), ("n_media", ctypes.c_int),
                ("material", ctypes.c_int * MAX_LAYERS),
                ("has_locals", ctypes.c_int * MAX_LAYERS)]


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
    magnet=(0, 0, 0)); a layer without density has no locals callback."""
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


def planar_callbacks(geometry):
    """The planar media of `geometry` behind the callback interface."""
    cb = Callbacks()
    _check(library().pumas_b200_planar_callbacks(
        ctypes.byref(geometry), ctypes.byref(cb)))
    return cb


def make_states(n, energy, charge=-1., direction=(0., 0., 1.), weight=1.,
                position=(0., 0., 0.)):
    s = np.zeros(n, dtype=STATE_DTYPE)
    s["charge"] = charge
    s["energy"] = energy
    s["weight"] = weight
    s["direction"] = np.asarray(direction, dtype="f8")
    s["position"] = np.asarray(position, dtype="f8")
    return s


class PumasError(RuntimeError):
    def __init__(self, rc, message):
        name = RETURN_NAMES[rc] if 0 <= rc < len(RETURN_NAMES) else str(rc)
        super().__init__("PUMAS_RETURN_%s: %s" % (name, message))
        self.rc = rc


_lib = None
_HANDLER_T = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_void_p,
                              ctypes.c_char_p)
_last_error = {"rc": 0, "message": ""}


@_HANDLER_T
def _handler(rc, caller, message):
    _last_error["rc"] = rc
    _last_error["message"] = (message or b"").decode(errors="replace")


def library():
    """Load libpumas_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBRARY):
        raise ImportError(
            "%s is missing: build it with `python -m pumas_b200.build` (nvcc, "
            "sm_100a). There is no CPU fallback." % LIBRARY)
    lib = ctypes.CDLL(LIBRARY)
    c_int_p = ctypes.POINTER(ctypes.c_int)
    c_double_p = ctypes.POINTER(ctypes.c_double)
    lib.pumas_physics_load.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
    lib.pumas_physics_dump.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.pumas_physics_destroy.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
    lib.pumas_physics_material_index.argtypes = [
        ctypes.c_void_p, ctypes.c_char_p, c_int_p]
    lib.pumas_physics_particle.argtypes = [
        ctypes.c_void_p, c_int_p, c_double_p, c_double_p]
    lib.pumas_physics_cutoff.restype = ctypes.c_double
    lib.pumas_physics_cutoff.argtypes = [ctypes.c_void_p]
    lib.pumas_context_create.argtypes = [
        ctypes.POINTER(ctypes.POINTER(ContextStruct)), ctypes.c_void_p,
        ctypes.c_int]
    lib.pumas_context_destroy.argtypes = [
        ctypes.POINTER(ctypes.POINTER(ContextStruct))]
    lib.pumas_context_random_seed_set.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.POINTER(ctypes.c_ulong)]
    lib.pumas_b200_context_geometry_set.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.POINTER(Geometry)]
    lib.pumas_b200_context_callbacks_set.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.POINTER(Callbacks)]
    lib.pumas_b200_planar_callbacks.argtypes = [
        ctypes.POINTER(Geometry), ctypes.POINTER(Callbacks)]
    lib.pumas_b200_planar_callbacks_free.argtypes = [ctypes.POINTER(Callbacks)]
    lib.pumas_b200_planar_callbacks_free.restype = None
    lib.pumas_b200_context_track_offset_set.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.c_ulonglong]
    lib.pumas_b200_context_stream_set.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.c_void_p]
    lib.pumas_b200_context_kernel_set.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.c_int]
    lib.pumas_context_transport_n.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.c_long, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]
    lib.pumas_b200_context_last_counters.argtypes = [
        ctypes.POINTER(ContextStruct), ctypes.c_long, ctypes.c_void_p,
        ctypes.c_void_p]
    lib.pumas_b200_context_last_kernel_time.argtypes = [
        ctypes.POINTER(ContextStruct), c_double_p, c_int_p]
    lib.pumas_b200_tally_scratch_bytes.restype = ctypes.c_long
    lib.pumas_b200_tally_scratch_bytes.argtypes = [ctypes.c_int]
    lib.pumas_b200_tally.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
        ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]
    lib.pumas_b200_fp64_peak.argtypes = [c_double_p]
    lib.pumas_recorder_create.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
    lib.pumas_recorder_clear.argtypes = [ctypes.c_void_p]
    lib.pumas_recorder_destroy.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
    lib.pumas_b200_recorder_export.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]
    lib.pumas_b200_context_transport_async.argtypes = [
        ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p]
    lib.pumas_b200_context_synchronize.argtypes = [ctypes.c_void_p]
    lib.pumas_b200_tally_accumulate.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
    lib.pumas_b200_tally_segments.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.pumas_b200_comm_unique_id.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.pumas_b200_comm_create.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
        ctypes.c_void_p, ctypes.c_int]
    lib.pumas_b200_comm_allreduce.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
    lib.pumas_b200_comm_destroy.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
    lib.pumas_error_handler_set.argtypes = [_HANDLER_T]
    lib.pumas_error_handler_set(_handler)
    _lib = lib
    return lib


def _check(rc):
    if rc != RETURN_SUCCESS:
        message = _last_error["message"] if _last_error["rc"] == rc else ""
        raise PumasError(rc, message)


def device_available():
    return bool(library().pumas_b200_device_available())


def fp64_peak():
    """Measured FP64 FMA peak of the current device, TFLOP/s."""
    t = ctypes.c_double()
    _check(library().pumas_b200_fp64_peak(ctypes.byref(t)))
    return t.value


def tally_scratch_bytes(n_bins):
    return library().pumas_b200_tally_scratch_bytes(n_bins)


def tally(states_ptr, events_ptr, n, event_mask, energy_min, energy_max,
          n_bins, out_ptr, scratch_ptr, stream=0):
    """pumas_b200_tally on device pointers (integers)."""
    _check(library().pumas_b200_tally(
        ctypes.c_void_p(states_ptr), ctypes.c_void_p(events_ptr), n,
        event_mask, energy_min, energy_max, n_bins, ctypes.c_void_p(out_ptr),
        ctypes.c_void_p(scratch_ptr), ctypes.c_void_p(stream)))


def tally_accumulate(total_ptr, batch_ptr, count, stream=0):
    """total[k] += batch[k] on the device (device pointers as integers)."""
    _check(library().pumas_b200_tally_accumulate(
        ctypes.c_void_p(total_ptr), ctypes.c_void_p(batch_ptr), count,
        ctypes.c_void_p(stream)))


def tally_segments(states_ptr, events_ptr, n_bins, per_bin, event_mask,
                   out_ptr, stream=0):
    """pumas_b200_tally_segments on device pointers (integers)."""
    _check(library().pumas_b200_tally_segments(
        ctypes.c_void_p(states_ptr), ctypes.c_void_p(events_ptr), n_bins,
        per_bin, event_mask, ctypes.c_void_p(out_ptr),
        ctypes.c_void_p(stream)))


class Comm:
    """The library's NCCL communicator (pumas_b200_comm_*): one rank per GPU.
    `exchange(bytes_or_None) -> bytes` hands the unique id made on rank 0 to
    every rank (e.g. a torch.distributed / MPI broadcast)."""

    def __init__(self, rank, world, exchange):
        lib = library()
        if not lib.pumas_b200_comm_available():
            raise RuntimeError("NCCL (libnccl.so.2) could not be loaded")
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _check(lib.pumas_b200_comm_unique_id(buf, 128))
        uid = exchange(bytes(buf.raw) if rank == 0 else None)
        buf = ctypes.create_string_buffer(bytes(uid), 128)
        self._ptr = ctypes.c_void_p()
        _check(lib.pumas_b200_comm_create(
            ctypes.byref(self._ptr), rank, world, buf, 128))
        self.rank, self.world = rank, world

    def allreduce(self, values_ptr, count, stream=0):
        """Sum `count` doubles at the device pointer over the ranks, in
        place, on `stream`."""
        _check(library().pumas_b200_comm_allreduce(
            self._ptr, ctypes.c_void_p(values_ptr), count,
            ctypes.c_void_p(stream)))

    def close(self):
        if self._ptr:
            library().pumas_b200_comm_destroy(ctypes.byref(self._ptr))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_libc = ctypes.CDLL(None)
_libc.fopen.restype = ctypes.c_void_p
_libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
_libc.fclose.argtypes = [ctypes.c_void_p]


class Physics:
    """struct pumas_physics loaded from a pumas_physics_dump file."""

    def __init__(self, dump_path):
        lib = library()
        self._handle = ctypes.c_void_p()
        stream = _libc.fopen(os.fsencode(dump_path), b"rb")
        if not stream:
            raise PumasError(17, "could not open `%s'" % dump_path)
        try:
            _check(lib.pumas_physics_load(ctypes.byref(self._handle), stream))
        finally:
            _libc.fclose(stream)

    def __del__(self):
        if getattr(self, "_handle", None) and _lib is not None:
            _lib.pumas_physics_destroy(ctypes.byref(self._handle))

    def dump(self, path):
        stream = _libc.fopen(os.fsencode(path), b"wb")
        if not stream:
            raise PumasError(17, "could not open `%s'" % path)
        try:
            _check(library().pumas_physics_dump(self._handle, stream))
        finally:
            _libc.fclose(stream)

    def material_index(self, name):
        i = ctypes.c_int(-1)
        _check(library().pumas_physics_material_index(
            self._handle, name.encode(), ctypes.byref(i)))
        return i.value

    def particle(self):
        kind = ctypes.c_int()
        ctau = ctypes.c_double()
        mass = ctypes.c_double()
        _check(library().pumas_physics_particle(
            self._handle, ctypes.byref(kind), ctypes.byref(ctau),
            ctypes.byref(mass)))
        return kind.value, ctau.value, mass.value

    def property(self, name, *args):
        """pumas_physics_property_<name>(physics, *args, &value)."""
        f = getattr(library(), "pumas_physics_property_" + name)
        value = ctypes.c_double()
        cargs = [ctypes.c_double(a) if isinstance(a, float) else ctypes.c_int(a)
                 for a in args]
        _check(f(self._handle, *cargs, ctypes.byref(value)))
        return value.value


class Context:
    """struct pumas_context: settings + device buffers of one stream."""

    def __init__(self, physics, seed=0):
        lib = library()
        self.physics = physics
        self._ptr = ctypes.POINTER(ContextStruct)()
        _check(lib.pumas_context_create(
            ctypes.byref(self._ptr), physics._handle, 0))
        self.seed(seed)

    def __del__(self):
        if getattr(self, "_ptr", None) and _lib is not None:
            _lib.pumas_context_destroy(ctypes.byref(self._ptr))

    @property
    def api(self):
        """The public struct pumas_context (mode, event, limit, accuracy)."""
        return self._ptr.contents

    def configure(self, settings):
        """Copy mode / event / limit / accuracy from another ContextStruct."""
        api = self.api
        api.mode.energy_loss = settings.mode.energy_loss
        api.mode.decay = settings.mode.decay
        api.mode.direction = settings.mode.direction
        api.mode.scattering = settings.mode.scattering
        api.event = settings.event
        api.limit.energy = settings.limit.energy
        api.limit.distance = settings.limit.distance
        api.limit.grammage = settings.limit.grammage
        api.limit.time = settings.limit.time
        api.accuracy = settings.accuracy

    def seed(self, seed):
        s = ctypes.c_ulong(seed)
        _check(library().pumas_context_random_seed_set(
            self._ptr, ctypes.byref(s)))

    def geometry(self, geometry):
        _check(library().pumas_b200_context_geometry_set(
            self._ptr, ctypes.byref(geometry)))

    def callbacks(self, callbacks):
        """`__device__` medium / locals functions (a Callbacks table, or None
        to go back to the descriptor)."""
        _check(library().pumas_b200_context_callbacks_set(
            self._ptr, None if callbacks is None else ctypes.byref(callbacks)))

    def track_offset(self, offset):
        _check(library().pumas_b200_context_track_offset_set(self._ptr, offset))

    def stream(self, cuda_stream):
        _check(library().pumas_b200_context_stream_set(
            self._ptr, ctypes.c_void_p(cuda_stream)))

    def kernel(self, kernel):
        """0: resident kernel (default), 1: monolithic kernel, 2: resident
        kernel, general instantiation only."""
        _check(library().pumas_b200_context_kernel_set(self._ptr, kernel))

    def transport(self, states, events=None, media=None):
        """pumas_context_transport_n on a host numpy array of STATE_DTYPE
        (updated in place). Returns (events, media)."""
        assert states.dtype == STATE_DTYPE and states.flags["C_CONTIGUOUS"]
        n = len(states)
        if events is None:
            events = np.zeros(n, dtype=np.int32)
        if media is None:
            media = np.zeros((n, 2), dtype=np.int32)
        _check(library().pumas_context_transport_n(
            self._ptr, n, states.ctypes.data, events.ctypes.data,
            None if media is False else media.ctypes.data))
        return events, media

    def transport_device(self, n, states_ptr, events_ptr=None, media_ptr=None):
        """pumas_context_transport_n on device pointers (integers)."""
        _check(library().pumas_context_transport_n(
            self._ptr, n, ctypes.c_void_p(states_ptr),
            ctypes.c_void_p(events_ptr), ctypes.c_void_p(media_ptr)))

    def record(self, period=1):
        """Attach a pumas_recorder of the given sampling period (pumas.h:
        432-458; `period` <= 0: medium changes and events only), or detach it
        (`period=None`)."""
        lib = library()
        if getattr(self, "_recorder", None):
            lib.pumas_recorder_destroy(ctypes.byref(self._recorder))
            self._recorder = None
            self.api.recorder = None
        if period is None:
            return
        self._recorder = ctypes.c_void_p()
        _check(lib.pumas_recorder_create(ctypes.byref(self._recorder), 0))
        # struct pumas_recorder: first (ptr), length (int), period (int)
        ctypes.c_int.from_address(self._recorder.value + 12).value = period
        self.api.recorder = self._recorder.value

    def frames(self, clear=True):
        """The frames recorded so far: (states, media, events) as arrays, in
        track order."""
        lib = library()
        n = lib.pumas_b200_recorder_export(
            self._ptr, self._recorder, 0, None, None, None)
        states = np.zeros(n, dtype=STATE_DTYPE)
        media = np.zeros(n, dtype=np.int32)
        events = np.zeros(n, dtype=np.int32)
        lib.pumas_b200_recorder_export(
            self._ptr, self._recorder, n, states.ctypes.data,
            media.ctypes.data, events.ctypes.data)
        if clear:
            lib.pumas_recorder_clear(self._recorder)
        return states, media, events

    def transport_async(self, n, states_ptr, events_ptr=None, media_ptr=None):
        """pumas_b200_context_transport_async: enqueue a batch (device
        pointers) on the context's stream and return."""
        _check(library().pumas_b200_context_transport_async(
            self._ptr, n, ctypes.c_void_p(states_ptr),
            ctypes.c_void_p(events_ptr), ctypes.c_void_p(media_ptr)))

    def synchronize(self):
        """Wait for the batches enqueued by transport_async; raises the first
        error one of their tracks met."""
        _check(library().pumas_b200_context_synchronize(self._ptr))

    def last_counters(self, n):
        nrand = np.zeros(n, dtype=np.uint32)
        nsteps = np.zeros(n, dtype=np.uint32)
        _check(library().pumas_b200_context_last_counters(
            self._ptr, n, nrand.ctypes.data, nsteps.ctypes.data))
        return nrand, nsteps

    def last_kernel_time(self):
        ms = ctypes.c_double()
        launches = ctypes.c_int()
        _check(library().pumas_b200_context_last_kernel_time(
            self._ptr, ctypes.byref(ms), ctypes.byref(launches)))
        return ms.value, launches.value
