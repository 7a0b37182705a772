#!/bin/bash
# pass V: dynamic tile scheduler + head-stream priority
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/tc_debug.py multi > gpurun_out/v_multi.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/v_multi.log | cut -c1-150
timeout -s KILL 300 python scripts/tc_debug.py pair > gpurun_out/v_pair.log 2>&1; echo "pair rc=$?"; tail -4 gpurun_out/v_pair.log | cut -c1-150
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/v_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/v_ops.log
timeout -s KILL 600 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/v_e2e.log 2>&1; echo "e2e rc=$?"; tail -3 gpurun_out/v_e2e.log
timeout -s KILL 240 python scripts/tc_debug.py epi > gpurun_out/v_epi.log 2>&1; head -5 gpurun_out/v_epi.log
for dyn in 1 0; do
  EDGECAPE_GEMM_DYNAMIC=$dyn timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/v_bench_dyn$dyn.log 2>&1; echo "bench dyn=$dyn rc=$?"
  tail -1 gpurun_out/v_bench_dyn$dyn.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'eager', round(d['eager_ms_per_step'],2), d['clocks']['reasons'])"
done
