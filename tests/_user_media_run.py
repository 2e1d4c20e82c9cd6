"""User `__device__` code against the reference: the medium / locals callbacks
of examples/user_media.h (the reference drives the same geometry through host
callbacks) and the DCS model of examples/user_dcs.h (registered with the
reference, which also made the tables). Run with PUMAS_B200_LIBRARY pointing to
the library built by `python -m pumas_b200.build --user examples/user_media.cu`;
prints one JSON line. usage: _user_media_run.py media|dcs n"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness
import pumas_b200 as pb
from workloads import dcs_models, isotropic, log_uniform

MAX_SHELLS = 8


class UserMedia(ctypes.Structure):
    """struct user_media (examples/user_media.h)."""
    _fields_ = [("n_shells", ctypes.c_int), ("reserved", ctypes.c_int),
                ("r", ctypes.c_double * (MAX_SHELLS + 1)),
                ("density", ctypes.c_double * MAX_SHELLS),
                ("gradient", ctypes.c_double * MAX_SHELLS),
                ("step_epsilon", ctypes.c_double)]


def shells(material):
    g = UserMedia()
    g.n_shells = 3
    for i, r in enumerate((0., 50., 120., 400.)):
        g.r[i] = r
    for i, (rho, a) in enumerate(((2.65e3, 1e-3), (1.0e3, 0.), (1.205, 2e-4))):
        g.density[i] = rho
        g.gradient[i] = a
    g.step_epsilon = 1e-7
    materials = (ctypes.c_int * MAX_SHELLS)(
        material("StandardRock"), material("Water"), material("Air"))
    return g, materials


def summary(c, so, sg, eo, eg, mo, mg, rc, steps):
    sync = ~c["diverged"]
    # The direction of a track that came to rest is ill conditioned (the last
    # steps scatter by large angles, scaled by step lengths that differ in the
    # last bits): counted apart. Everything else must hold 1e-9.
    stopped = (so["energy"] <= 0.) & (sg["energy"] <= 0.)
    masked_o, masked_g = so.copy(), sg.copy()
    masked_g["direction"][stopped] = masked_o["direction"][stopped]
    c2 = harness.compare_states(masked_o, masked_g, rtol=1e-9)
    d = np.abs(so["direction"] - sg["direction"]).max(axis=1)
    return dict(
        off_but_rest_direction=int(((c2["worst"] > 1e-9) & sync).sum()),
        max_rel_but_rest_direction=float(c2["worst"][sync].max()),
        rest_direction_max=float(d[stopped].max()) if stopped.any() else 0.,
        rc=int(rc), n=int(c["n"]), match=int(c["n_match"]),
        bitexact=int(c["n_bitexact"]), diverged=int(c["n_diverged"]),
        off_in_sync=int(((c["worst"] > 1e-9) & sync).sum()),
        max_rel=float(c["max_rel"]),
        media_equal=bool((mo[sync] == mg[sync]).all()),
        exits=int((mg[:, 1] < 0).sum()),
        media_changes=int((eg & pb.EVENT_MEDIUM != 0).sum()),
        steps_per_track=float(steps.mean()))


def main_dcs(n):
    """Tables made by the reference with the user's bremsstrahlung model
    (tests/golden/rock.muon.usr.pumas.gz); hard losses evaluate the model at
    run time (pumas.c:11279-11346)."""
    lib = pb.library()
    out = {}
    # before the registration the dump is refused, as the reference does
    try:
        pb.Physics(harness.dump_path("rock.muon.usr"))
        out["unregistered_rc"] = 0
    except pb.PumasError as e:
        out["unregistered_rc"] = e.rc
    assert lib.user_dcs_register() == 0
    physics = pb.Physics(harness.dump_path("rock.muon.usr"))
    lib_o = ctypes.CDLL(harness.ORACLE_SO)
    assert lib_o.oracle_register_user_dcs() == 0
    oracle = harness.Oracle("rock.muon.usr")
    for name, direction in (("forward", pb.MODE_FORWARD),
                            ("backward", pb.MODE_BACKWARD)):
        w = dcs_models(oracle.material, n, direction=direction)
        so, eo, mo, ro, rc, _ = oracle.transport(
            w["settings"], w["geometry"], w["states"], seed=13, offset=900,
            threads=os.cpu_count() or 1)
        ctx = pb.Context(physics, seed=13)
        ctx.configure(w["settings"])
        ctx.geometry(w["geometry"])
        ctx.track_offset(900)
        sg = w["states"].copy()
        eg, mg = ctx.transport(sg)
        rg, steps = ctx.last_counters(n)
        c = harness.compare_states(so, sg, eo, eg, ro, rg, rtol=1e-9)
        out[name] = summary(c, so, sg, eo, eg, mo, mg, rc, steps)
    print(json.dumps(out))


def main(n):
    oracle = harness.Oracle("materials.muon")
    physics = pb.Physics(harness.dump_path("materials.muon"))
    g, materials = shells(oracle.material)
    lib = pb.library()
    cb = pb.Callbacks()
    lib.user_media_callbacks.argtypes = [
        ctypes.POINTER(UserMedia), ctypes.POINTER(ctypes.c_int),
        ctypes.POINTER(pb.Callbacks)]
    assert lib.user_media_callbacks(
        ctypes.byref(g), materials, ctypes.byref(cb)) == 0
    oracle.lib.oracle_set_user_media.argtypes = [
        ctypes.POINTER(UserMedia), ctypes.POINTER(ctypes.c_int)]
    oracle.lib.oracle_set_user_media(ctypes.byref(g), materials)

    out = {}
    cases = {
        "forward": dict(
            settings=harness.make_settings(pb.MODE_STRAGGLED, pb.MODE_MIXED),
            states=pb.make_states(n, log_uniform(n, 0.5, 200., 1),
                                  direction=isotropic(n, 2),
                                  position=(10., 5., -20.))),
        "backward": dict(
            settings=harness.make_settings(
                pb.MODE_STRAGGLED, pb.MODE_MIXED, direction=pb.MODE_BACKWARD,
                event=pb.EVENT_LIMIT_ENERGY, limit_energy=1e3),
            states=pb.make_states(n, log_uniform(n, 0.5, 50., 3),
                                  direction=isotropic(n, 4),
                                  position=(10., 5., -20.))),
        "csda": dict(
            settings=harness.make_settings(pb.MODE_CSDA, pb.MODE_DISABLED),
            states=pb.make_states(n, log_uniform(n, 0.5, 200., 5),
                                  direction=isotropic(n, 6),
                                  position=(0., 60., 0.))),
    }
    dummy = pb.uniform_geometry(0)
    for name, w in cases.items():
        so, eo, mo, ro, rc, _ = oracle.transport(
            w["settings"], dummy, w["states"], seed=11, offset=500,
            threads=os.cpu_count() or 1)
        ctx = pb.Context(physics, seed=11)
        ctx.configure(w["settings"])
        ctx.callbacks(cb)
        ctx.track_offset(500)
        sg = w["states"].copy()
        eg, mg = ctx.transport(sg)
        rg, steps = ctx.last_counters(n)
        c = harness.compare_states(so, sg, eo, eg, ro, rg, rtol=1e-9)
        out[name] = summary(c, so, sg, eo, eg, mo, mg, rc, steps)
    oracle.lib.oracle_set_user_media(None, None)
    print(json.dumps(out))


if __name__ == "__main__":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    if (len(sys.argv) > 1) and (sys.argv[1] == "dcs"):
        main_dcs(n)
    else:
        main(n)
