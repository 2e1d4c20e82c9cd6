"""Profiling / timing driver: one workload through the C ABI.
usage: _gpu_prof.py workload n reps [kernel]"""
import sys
sys.path.insert(0, 'tests')
import harness, pumas_b200 as pb
from workloads import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "headline"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dump, builder, kw = WORKLOADS[name]
ph = pb.Physics(harness.dump_path(dump))
w = builder(ph.material_index, n, **kw)
ctx = pb.Context(ph, seed=1); ctx.configure(w["settings"]); ctx.geometry(w["geometry"]); ctx.kernel(kernel)
for rep in range(reps):
    s = w["states"].copy()
    ctx.track_offset(0)
    ctx.transport(s)
    ms, _ = ctx.last_kernel_time()
    print(f"{name} n={n} k={kernel}: kernel {ms:.2f} ms -> {n/ms*1e3:.3e} tracks/s", flush=True)
