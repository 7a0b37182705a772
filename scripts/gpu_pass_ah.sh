#!/bin/bash
# pass AH: validation after the one-kernel GCN became the default -- full GPU suite, smoke, bench, ncu of the GCN kernels
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu > gpurun_out/ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ah_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ah_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/ah_bench.log > gpurun_out/ah_bench.json
python -c "import json; d=json.load(open('gpurun_out/ah_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['reasons'], d['gpu_launches'], json.dumps(d['north_star_kernels']), d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gcn_fused -s 3 -c 2 -f -o gpurun_out/prof_ah_gcn_fused python scripts/gcn_once.py > gpurun_out/ah_ncu.log 2>&1; echo "ncu rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none -k regex:"gcn_aggregate|gemm_f16x3" -s 6 -c 4 -f -o gpurun_out/prof_ah_gcn_two_kernel python scripts/gcn_once.py --two-kernel > gpurun_out/ah_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout -s KILL 200 python scripts/gcn_bench.py > gpurun_out/ah_gcn_bench.jsonl 2>&1; echo "gcn bench rc=$?"
