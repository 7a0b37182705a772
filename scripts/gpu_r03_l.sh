#!/bin/bash
# round 2, session 2, pass L: K = 200 fused GCN -- micro-benchmark vs the two-launch path, goldens, configs[4] slice bench
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/gcn_bench.py --k200 --fused-only 2>&1 | grep '"K": 200' | cut -c1-175 > gpurun_out/r03l_gcn_bench_k200.jsonl; cat gpurun_out/r03l_gcn_bench_k200.jsonl
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x > gpurun_out/r03l_pytest_e2e.log 2>&1; echo "pytest e2e rc=$?"; tail -3 gpurun_out/r03l_pytest_e2e.log
for f in 2 0; do
EDGECAPE_GCN_FUSED=$f timeout -s KILL 600 python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/r03l_bench_c5_fused$f.log 2>&1; echo "bench c5 fused=$f rc=$?"
tail -1 gpurun_out/r03l_bench_c5_fused$f.log > gpurun_out/r03l_bench_c5_fused$f.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03l_bench_c5_fused$f.json'))
    print('c5 fused=$f', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'], 'parity', d.get('parity',{}).get('max_rel_err'), d.get('parity',{}).get('argmax_equal'))
except Exception as e:
    print('bench parse failed', e)
PY
done
