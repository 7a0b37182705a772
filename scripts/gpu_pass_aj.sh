#!/bin/bash
# pass AJ: final validation of round 1 -- full GPU suite (with the fused-GCN edge shapes), smoke, default bench, reference arm
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/aj_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/aj_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 400 python bench.py > gpurun_out/aj_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/aj_bench.log > gpurun_out/aj_bench.json
python -c "import json; d=json.load(open('gpurun_out/aj_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'], round(d['north_star_kernels']['gcn']['us'],2), round(d['roofline']['frac'],3), d['cpu_baseline']['value'])"
