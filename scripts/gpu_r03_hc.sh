#!/bin/bash
# round 2, session 2, last experiment: the head graph's persistent kernels narrowed to n CTAs (EDGECAPE_HEAD_CTAS)
mkdir -p gpurun_out
for n in 0 48 96; do
  EDGECAPE_HEAD_CTAS=$n timeout -s KILL 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 2>/dev/null | tail -1 > gpurun_out/r03hc_bench_$n.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03hc_bench_$n.json'))
    print('head ctas $n:', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
except Exception as e:
    print('head ctas $n: bench parse failed', e)
PY
done
EDGECAPE_HEAD_CTAS=48 timeout -s KILL 60 python scripts/overlap_probe.py 2>&1 | tail -1
EDGECAPE_HEAD_CTAS=48 timeout -s KILL 150 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x 2>&1 | tail -2
