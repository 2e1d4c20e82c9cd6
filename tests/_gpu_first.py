import sys, time, os
sys.path.insert(0, 'tests')
import numpy as np
import harness, pumas_b200 as pb
from workloads import WORKLOADS
print("device:", pb.device_available())
o = harness.Oracle(); ph = pb.Physics(harness.dump_path())
for name, n in (("csda_water", 10_000_000), ("straight", 200_000), ("headline", 200_000), ("headline", 1_000_000)):
    dump, builder, kw = WORKLOADS[name]
    w = builder(o.material, n, **kw)
    ctx = pb.Context(ph, seed=1); ctx.configure(w["settings"]); ctx.geometry(w["geometry"])
    for rep in range(2):
        s = w["states"].copy()
        t0 = time.time(); ctx.transport(s); t1 = time.time()
        ms, _ = ctx.last_kernel_time()
        nr, ns = ctx.last_counters(n)
        print(f"{name} n={n}: kernel {ms:.1f} ms -> {n/ms*1e3:.3e} tracks/s ; e2e {t1-t0:.3f}s -> {n/(t1-t0):.3e} ; steps/track {ns.mean():.1f}", flush=True)
# cpu reference speed
w = WORKLOADS["headline"][1](o.material, 4000)
r = o.transport(w["settings"], w["geometry"], w["states"], threads=os.cpu_count())
print("cpu ref headline:", 4000/r[5], "tracks/s on", os.cpu_count(), "threads")
