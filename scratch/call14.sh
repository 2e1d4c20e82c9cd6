#!/bin/bash
python - <<'PY'
import sys, os
sys.path.insert(0,'tests')
import numpy as np, torch, harness, pumas_b200 as pb
from workloads import WORKLOADS
name,n="straight_mixed",1<<21
dump, builder, kw = WORKLOADS[name]
ph = pb.Physics(harness.dump_path(dump))
w = builder(ph.material_index, n, **kw)
for kernel in (0,3):
    ctx = pb.Context(ph, seed=2026); ctx.configure(w["settings"]); ctx.geometry(w["geometry"]); ctx.kernel(kernel)
    for k in (0,1,2):
        s = w["states"].copy()
        ctx.track_offset(k*n)
        ev, md = ctx.transport(s)
        ms,_ = ctx.last_kernel_time()
        nr, ns = ctx.last_counters(n)
        i = int(np.argmax(ns))
        print("kernel",kernel,"k",k,"ms %.1f"%ms,"steps mean %.1f max %d (track %d: E %.4g event %d nrand %d)"%(ns.mean(), ns.max(), i, s["energy"][i], ev[i], nr[i]), "top5", np.sort(ns)[-5:], flush=True)
PY
