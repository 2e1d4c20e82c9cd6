#!/bin/bash
run() { timeout 300 python tests/_gpu_prof.py $1 $2 2 0 2>&1 | tail -1 | sed 's/.*-> //'; }
for wl in straight_mixed straight rock_3km; do
for d in 12 16 21 28 36 48; do
  echo "$wl service_ctas=$d: $(PUMAS_B200_DEL_CTAS=$d run $wl 2000000)"
done; done
echo "csda_water 8M: $(run csda_water 8388608)"
python - <<'PY'
import sys, os
sys.path.insert(0,'tests')
import numpy as np, torch, harness, pumas_b200 as pb
from workloads import WORKLOADS
for name,n in (("csda_water",1<<23),("straight_mixed",1<<21)):
    dump, builder, kw = WORKLOADS[name]
    ph = pb.Physics(harness.dump_path(dump))
    w = builder(ph.material_index, n, **kw)
    ctx = pb.Context(ph, seed=1); ctx.configure(w["settings"]); ctx.geometry(w["geometry"])
    ctx.stream(torch.cuda.current_stream().cuda_stream)
    pristine = torch.from_numpy(w["states"].view(np.uint8).reshape(n,104).copy()).cuda()
    work = torch.empty_like(pristine); events = torch.zeros(n, dtype=torch.int32, device="cuda")
    for k in range(4):
        work.copy_(pristine, non_blocking=True)
        ctx.track_offset(k*n)
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.transport_async(n, work.data_ptr(), events.data_ptr()); e1.record(); ctx.synchronize()
        print(name, k, "events ms", e0.elapsed_time(e1), "internal ms", ctx.last_kernel_time())
PY
