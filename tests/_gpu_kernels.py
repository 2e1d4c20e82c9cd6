"""Kernel A/B on the GPU: for each workload run the kernels given (default:
resident 0 and monolithic 1), compare them bit for bit and print the timings.
usage: _gpu_kernels.py [-k 0,1,2] workload:n ..."""
import sys, time, os
sys.path.insert(0, 'tests')
import numpy as np
import harness, pumas_b200 as pb
from workloads import WORKLOADS
args = sys.argv[1:]
kernels = (0, 1)
if args and args[0] == "-k":
    kernels = tuple(int(k) for k in args[1].split(","))
    args = args[2:]
runs = [(a.split(":")[0], int(a.split(":")[1])) for a in args] or [("headline", 200000)]
phys = {}
for name, n in runs:
    dump, builder, kw = WORKLOADS[name]
    if dump not in phys:
        phys[dump] = pb.Physics(harness.dump_path(dump))
    ph = phys[dump]
    w = builder(ph.material_index, n, **kw)
    ref = None
    for kernel in kernels:
        ctx = pb.Context(ph, seed=1); ctx.configure(w["settings"]); ctx.geometry(w["geometry"]); ctx.kernel(kernel)
        for rep in range(2):
            s = w["states"].copy()
            ctx.track_offset(0)
            ev, md = ctx.transport(s)
            ms, _ = ctx.last_kernel_time()
            nr, ns = ctx.last_counters(n)
        print(f"{name} n={n} kernel={kernel}: {ms:.1f} ms -> {n/ms*1e3:.3e} tracks/s ; steps/track {ns.mean():.1f} ; {ns.sum()/ms*1e3:.3e} steps/s", flush=True)
        if ref is None:
            ref = (s, ev, nr, md)
        else:
            bad = np.zeros(n, bool)
            for f in harness.STATE_FIELDS:
                a, b = ref[0][f], s[f]
                d = (a != b) & ~(np.isnan(a) & np.isnan(b))
                bad |= d.reshape(n, -1).any(axis=1)
            bad |= (ref[1] != ev) | (ref[2] != nr) | (ref[3].reshape(n, -1) != md.reshape(n, -1)).any(axis=1)
            print(f"   kernel {kernel} == kernel {kernels[0]} bit for bit: {not bad.any()} ({bad.sum()} tracks differ)", flush=True)
            if bad.any():
                i = np.flatnonzero(bad)[0]
                print("   first:", i, ref[0][i], s[i], ref[1][i], ev[i], ref[2][i], nr[i])
