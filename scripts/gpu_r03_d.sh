#!/bin/bash
# round 2, session 2, pass D: pipeline depth 2 vs 3 (resident and end to end), same box, alternating
mkdir -p gpurun_out
for d in 2 3 2 3 4; do
  timeout -s KILL 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustained-seconds 0 --depth $d 2>/dev/null | tail -1 > gpurun_out/r03d_bench_depth$d.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03d_bench_depth$d.json'))
    print('depth $d', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3))
except Exception as e:
    print('depth $d bench parse failed', e)
PY
done
