#!/bin/bash
# round 2, session 2, pass E: attention kernels skip the softmax work of query-row quarters beyond Lq
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_e2e_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 2>/dev/null | tail -1 > gpurun_out/r03e_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03e_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'])
    print('nsk', {k:(v['us'],v['frac'], v.get('issued_frac')) for k,v in d['north_star_kernels'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
