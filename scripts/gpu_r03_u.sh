#!/bin/bash
# round 2, session 2, pass U: K-split of the tail tiles (opt-in) at the level of the step: bench off / on
mkdir -p gpurun_out
for k in 0 1 0 1; do
EDGECAPE_GEMM_KSPLIT=$k timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/r03u_bench_ksplit$k.log 2>&1
tail -1 gpurun_out/r03u_bench_ksplit$k.log > gpurun_out/r03u_bench_ksplit$k.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03u_bench_ksplit$k.json'))
    print('ksplit=$k', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], 'gemm ms/step', round(d['roofline']['kernel_ms_per_step'],3), round(d['roofline']['all_gemm_ms_per_step'],3))
except Exception as e:
    print('bench parse failed', e)
PY
done
