timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_resident -c 1 -o gpurun_out/prof_r1_v6 python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/prof_v6.log 2>&1
tail -2 gpurun_out/prof_v6.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --cpu-sample 2000 > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-300
echo "== above_grid_mixed detail"
timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0,'tests')
import numpy as np, harness, pumas_b200 as pb
from workloads import WORKLOADS
dump, builder, kw = WORKLOADS["above_grid_mixed"]
oracle = harness.Oracle(dump); ph = pb.Physics(harness.dump_path(dump))
n=20000
w = builder(oracle.material, n, **kw)
so, eo, mo, ro, rc, _ = oracle.transport(w["settings"], w["geometry"], w["states"], seed=7, offset=1000, threads=16)
for k in (0,1):
    ctx = pb.Context(ph, seed=7); ctx.configure(w["settings"]); ctx.geometry(w["geometry"]); ctx.track_offset(1000); ctx.kernel(k)
    s = w["states"].copy(); ev, md = ctx.transport(s); nr, ns = ctx.last_counters(n)
    c = harness.compare_states(so, s, eo, ev, ro, nr, rtol=1e-9)
    print("kernel",k,"match",c["n_match"],"diverged",c["n_diverged"],"max_rel",c["max_rel"])
    rel = np.zeros(n)
    for f in ("energy","distance","grammage","time","weight"):
        a=so[f]; b=s[f]; d=np.abs(a-b)/np.maximum(np.abs(a),1e-300); rel=np.maximum(rel,np.where(np.isfinite(d),d,0))
    bad=np.argsort(-rel)[:4]
    for i in bad: print("  track",i,"rel",rel[i],"Ki",w["states"]["energy"][i],"Kf ref",so["energy"][i],"gpu",s["energy"][i],"w ref",so["weight"][i],"gpu",s["weight"][i],"ev",eo[i],ev[i],"nrand",ro[i],nr[i])
PY
