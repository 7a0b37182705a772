#!/bin/bash
# tcgen05 GEMM bring-up: staged checks (each under its own timeout so a hung kernel cannot eat the box), then tests + bench
mkdir -p gpurun_out
for st in tiny k2 multi bench; do
  timeout -s KILL 120 python scripts/tc_debug.py $st > gpurun_out/tc_$st.log 2>&1
  echo "stage $st exit $?" >> gpurun_out/tc_$st.log
  tail -4 gpurun_out/tc_$st.log
done
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
