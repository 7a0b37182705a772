#!/bin/bash
# pass T: software-pipelined batches (async API), SM partition sweep
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?"
tail -5 gpurun_out/t_e2e.log
for ctas in 0 140 132 124 112; do
  EDGECAPE_GEMM_CTAS=$ctas timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/t_bench_$ctas.log 2>&1; echo "bench ctas=$ctas rc=$?"
  tail -1 gpurun_out/t_bench_$ctas.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'eager', round(d['eager_ms_per_step'],2), d['clocks'])"
done
