#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 --tracks 4194304 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c13_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['config']['extra'].items(): print(k, v.get('tracks_per_s'), v.get('launch_ms'), v.get('roofline_frac'), v.get('error'))
PY
