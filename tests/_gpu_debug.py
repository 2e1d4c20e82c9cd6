import sys, os
sys.path.insert(0, 'tests')
import numpy as np
import harness, pumas_b200 as pb
from workloads import WORKLOADS
names = sys.argv[1:] or ["decay_randomised", "no_loss", "mixed_loss_bwd", "geometry_detailed"]
cache = {}
for name in names:
    dump, builder, kw = WORKLOADS[name]
    if dump not in cache:
        cache[dump] = (harness.Oracle(dump), pb.Physics(harness.dump_path(dump)))
    o, ph = cache[dump]
    n = 256
    w = builder(o.material, n, **kw)
    so, eo, mo, ro, rc, _ = o.transport(w["settings"], w["geometry"], w["states"], seed=7, offset=1000, threads=4)
    res = {}
    for kernel in (0, 1):
        ctx = pb.Context(ph, seed=7); ctx.configure(w["settings"]); ctx.geometry(w["geometry"]); ctx.kernel(kernel); ctx.track_offset(1000)
        s = w["states"].copy(); ev, md = ctx.transport(s); nr, ns = ctx.last_counters(n)
        res[kernel] = (s, ev, md, nr, ns)
        c = harness.compare_states(so, s, eo, ev, ro, nr)
        print(f"{name} kernel={kernel}: match={c['n_match']} diverged={c['n_diverged']} max_rel={c['max_rel']:.3g}")
    s, ev, md, nr, ns = res[1]
    bad = np.nonzero((ro != nr) | (eo != ev))[0][:6]
    for i in bad:
        print(f"  track {i}: K0={w['states']['energy'][i]:.6g} ref: ev={eo[i]} nrand={ro[i]} K={so['energy'][i]:.17g} d={so['distance'][i]:.17g} t={so['time'][i]:.17g} w={so['weight'][i]:.17g}")
        print(f"            gpu: ev={ev[i]} nrand={nr[i]} K={s['energy'][i]:.17g} d={s['distance'][i]:.17g} t={s['time'][i]:.17g} w={s['weight'][i]:.17g} steps={ns[i]}")
    s0 = res[0][0]
    diff01 = [i for i in range(n) if any(not np.array_equal(s0[f][i], s[f][i]) for f in harness.STATE_FIELDS) or res[0][3][i] != nr[i]]
    print("  wavefront vs monolithic differing tracks:", diff01[:10], len(diff01))
    for i in diff01[:3]:
        print(f"     k0: ev={res[0][1][i]} nrand={res[0][3][i]} K={s0['energy'][i]:.17g} pos={s0['position'][i]} steps={res[0][4][i]}")
        print(f"     k1: ev={ev[i]} nrand={nr[i]} K={s['energy'][i]:.17g} pos={s['position'][i]} steps={ns[i]}")
