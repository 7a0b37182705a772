#!/bin/bash
# round 2, session 2, pass F: project-first fused GCN as the default -- GCN tests / trace / bench, the end-to-end goldens, smoke, default bench
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "gcn" > gpurun_out/r03f_pytest_gcn.log 2>&1; echo "pytest gcn rc=$?"; tail -2 gpurun_out/r03f_pytest_gcn.log
timeout -s KILL 200 python scripts/gcn2_trace.py --experiments > gpurun_out/r03f_gcn2_trace.log 2>&1; echo "trace rc=$?"; head -26 gpurun_out/r03f_gcn2_trace.log
timeout -s KILL 300 python scripts/gcn_bench.py --fused-only > gpurun_out/r03f_gcn_bench.jsonl 2>&1; echo "bench rc=$?"; cut -c1-170 gpurun_out/r03f_gcn_bench.jsonl
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x > gpurun_out/r03f_pytest_e2e.log 2>&1; echo "pytest e2e rc=$?"; tail -3 gpurun_out/r03f_pytest_e2e.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py > gpurun_out/r03f_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r03f_bench.log > gpurun_out/r03f_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03f_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','frac_of_sustained_peak','kernel_ms_per_step')})
    print('parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'], 'cpu', d['cpu_baseline']['value'])
    print('sustained', d['sustained']['value'], d['sustained']['clocks'])
    print('nsk', {k:(v['us'],v['frac']) for k,v in d['north_star_kernels'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
