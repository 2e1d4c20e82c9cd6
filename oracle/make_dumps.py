"""Generate the physics dumps used as fixtures (run by `make -C oracle dumps`).

The dumps are produced by the reference's own pumas_physics_create +
pumas_physics_dump (src/pumas.c:1916, 2167) through oracle/_ref. They are the
interchange format between the reference and the GPU library:
both sides load the same bytes. Outputs (gzip, committed):
    tests/golden/materials.muon.pumas.gz   examples/data/materials.xml, muon
    tests/golden/ice.tau.pumas.gz          oracle/data/ice.xml, tau
    tests/golden/rock.muon.kkb.pumas.gz    oracle/data/rock.xml, muon, models
                                           KKP / KKP / BBKS
    tests/golden/rock.muon.abm.pumas.gz    oracle/data/rock.xml, muon, models
                                           ABB / SSR / BM
    tests/golden/rock.muon.usr.pumas.gz    oracle/data/rock.xml, muon, the user
                                           registered bremsstrahlung model of
                                           examples/user_dcs.h (BHS) / SSR /
                                           DRSS
"""
import ctypes
import gzip
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libpumas_oracle.so"))
lib.oracle_physics_create_and_dump.argtypes = [
    ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]

lib.oracle_physics_create_and_dump_models.argtypes = [
    ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p,
    ctypes.c_char_p, ctypes.c_char_p]

JOBS = [
    (0, "examples/materials.xml", "materials.muon.pumas", None),
    (1, "ice/ice.xml", "ice.tau.pumas", None),
    (0, "rock/rock.xml", "rock.muon.kkb.pumas", (b"KKP", b"KKP", b"BBKS")),
    (0, "rock/rock.xml", "rock.muon.abm.pumas", (b"ABB", b"SSR", b"BM")),
    (0, "rock/rock.xml", "rock.muon.usr.pumas", (b"BHS", b"SSR", b"DRSS")),
]
assert lib.oracle_register_user_dcs() == 0
if len(sys.argv) > 1:
    JOBS = [j for j in JOBS if j[2] in sys.argv[1:]]

for particle, mdf, name, models in JOBS:
    mdf_path = os.path.join(HERE, "_ref", "data", mdf)
    raw = os.path.join(HERE, "_ref", "data", name)
    if models is None:
        rc = lib.oracle_physics_create_and_dump(
            particle, mdf_path.encode(), None, raw.encode())
    else:
        rc = lib.oracle_physics_create_and_dump_models(
            particle, mdf_path.encode(), raw.encode(), *models)
    if rc != 0:
        sys.exit(f"pumas_physics_create failed for {mdf} (rc={rc})")
    out = os.path.join(HERE, "..", "tests", "golden", name + ".gz")
    with open(raw, "rb") as f, gzip.GzipFile(out, "wb", mtime=0) as g:
        g.write(f.read())
    print(f"{name}: {os.path.getsize(raw)} B -> {os.path.getsize(out)} B gz")
