#!/bin/bash
# pass U: support de-duplication, full GPU suite, bench with the informational de-dup leg
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -x -q -m gpu > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/u_pytest.log
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dedup-demo > gpurun_out/u_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/u_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d.get('support_dedup_demo'))"
