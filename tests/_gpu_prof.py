"""Profiling driver (ncu): one headline batch through the C ABI."""
import sys
sys.path.insert(0, 'tests')
import harness, pumas_b200 as pb
from workloads import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "headline"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dump, builder, kw = WORKLOADS[name]
ph = pb.Physics(harness.dump_path(dump))
w = builder(ph.material_index, n, **kw)
ctx = pb.Context(ph, seed=1); ctx.configure(w["settings"]); ctx.geometry(w["geometry"])
for rep in range(reps):
    s = w["states"].copy()
    ctx.transport(s)
    ms, _ = ctx.last_kernel_time()
    print(f"{name} n={n}: kernel {ms:.2f} ms -> {n/ms*1e3:.3e} tracks/s", flush=True)
