"""Variant A/B on the GPU: time one workload with the library named by
PUMAS_B200_LIBRARY and print a digest of the per-track results (states, events,
variates drawn) so that builds can be compared bit for bit across processes.
usage: _gpu_variant.py [workload] [n] [reps]"""
import hashlib
import sys
sys.path.insert(0, 'tests')
import numpy as np
import harness, pumas_b200 as pb
from workloads import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "headline"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 600000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dump, builder, kw = WORKLOADS[name]
ph = pb.Physics(harness.dump_path(dump))
w = builder(ph.material_index, n, **kw)
ctx = pb.Context(ph, seed=1); ctx.configure(w["settings"]); ctx.geometry(w["geometry"])
best = 1e30
for rep in range(reps):
    s = w["states"].copy()
    ctx.track_offset(0)
    ev, md = ctx.transport(s)
    ms, _ = ctx.last_kernel_time()
    best = min(best, ms)
nr, ns = ctx.last_counters(n)
h = hashlib.sha1()
for a in (s, ev, nr, md):
    h.update(np.ascontiguousarray(a).tobytes())
import os
print(f"{os.path.basename(pb.LIBRARY):28s} {name} n={n}: {best:8.2f} ms -> {n/best*1e3:.4e} tracks/s  steps/track {ns.mean():.1f}  digest {h.hexdigest()[:16]}", flush=True)
