#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/w_e2e.log 2>&1; echo "e2e rc=$?"; tail -2 gpurun_out/w_e2e.log
for dyn in 1 0; do
EDGECAPE_GEMM_DYNAMIC=$dyn timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/w_bench$dyn.log 2>&1; echo "bench dyn=$dyn rc=$?"
tail -1 gpurun_out/w_bench$dyn.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'eager', round(d['eager_ms_per_step'],2), d['clocks']['reasons'])"
done
