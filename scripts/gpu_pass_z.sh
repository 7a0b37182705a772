#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/z_pytest.log
timeout -s KILL 200 python scripts/overlap_probe.py 2>&1 | tail -1
timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/z_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/z_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['reasons'])"
