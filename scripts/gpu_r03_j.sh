#!/bin/bash
# round 2, session 2, pass J: programmatic dependent launch inside the captured graphs -- none / backbone only / head only / both
mkdir -p gpurun_out
for m in none vit all none vit; do
EDGECAPE_PDL_GRAPHS=$m timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/r03j_bench_pdl_$m.log 2>&1; echo "bench pdl=$m rc=$?"
tail -1 gpurun_out/r03j_bench_pdl_$m.log > gpurun_out/r03j_bench_pdl_$m.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03j_bench_pdl_$m.json'))
    print('pdl=$m', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], 'parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'])
except Exception as e:
    print('bench parse failed', e)
PY
done
