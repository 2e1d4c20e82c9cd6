"""Generate tests/golden/tracks.npz: final states of the REFERENCE (oracle/_ref,
the unmodified src/pumas.c) for 48 tracks of every workload of
tests/workloads.py, Philox seed 21, track offset 4096. Run in the build
container (needs /root/reference to have built oracle/_ref):
    python tests/golden/make_golden.py
The fixture lets the GPU box (no /root/reference) and any later checkout verify
parity against recorded reference outputs as well as against the live oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness  # noqa: E402
from workloads import WORKLOADS  # noqa: E402

N, SEED, OFFSET = 48, 21, 4096
out = {}
cache = {}
for name, (dump, builder, kwargs) in sorted(WORKLOADS.items()):
    if dump not in cache:
        cache[dump] = harness.Oracle(dump)
    oracle = cache[dump]
    w = builder(oracle.material, N, **kwargs)
    s, e, m, r, rc, _ = oracle.transport(
        w["settings"], w["geometry"], w["states"], seed=SEED, offset=OFFSET)
    assert rc == 0
    out[name + ".states"] = s
    out[name + ".events"] = e
    out[name + ".media"] = m
    out[name + ".nrand"] = r
np.savez_compressed(os.path.join(HERE, "tracks.npz"), **out)
print("wrote", os.path.join(HERE, "tracks.npz"))
