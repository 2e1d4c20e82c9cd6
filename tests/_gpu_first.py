import sys, time, os
sys.path.insert(0, 'tests')
import numpy as np
import harness, pumas_b200 as pb
from workloads import WORKLOADS
print("device:", pb.device_available())
ph = pb.Physics(harness.dump_path())
runs = [("csda_water", 4_000_000), ("straight", 400_000), ("headline", 400_000), ("rock_3km", 200_000), ("geometry_detailed", 400_000)]
if len(sys.argv) > 1:
    runs = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
for name, n in runs:
    dump, builder, kw = WORKLOADS[name]
    w = builder(ph.material_index, n, **kw)
    ref = None
    for kernel in (0, 1):
        ctx = pb.Context(ph, seed=1); ctx.configure(w["settings"]); ctx.geometry(w["geometry"]); ctx.kernel(kernel)
        for rep in range(2):
            s = w["states"].copy()
            ctx.track_offset(0)
            t0 = time.time(); ev, md = ctx.transport(s); t1 = time.time()
            ms, _ = ctx.last_kernel_time()
            nr, ns = ctx.last_counters(n)
        print(f"{name} n={n} kernel={kernel}: {ms:.1f} ms -> {n/ms*1e3:.3e} tracks/s ; e2e {n/(t1-t0):.3e} ; steps/track {ns.mean():.1f} ; {ns.sum()/ms*1e3:.3e} steps/s", flush=True)
        if ref is None:
            ref = (s, ev, nr)
        else:
            same = all(np.array_equal(ref[0][f], s[f]) for f in harness.STATE_FIELDS) and np.array_equal(ref[1], ev) and np.array_equal(ref[2], nr)
            print(f"   wavefront == monolithic bit for bit: {same}", flush=True)
