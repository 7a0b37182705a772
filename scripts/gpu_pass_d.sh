#!/bin/bash
# pass D: tcgen05 attention bring-up, then the whole suite + bench
mkdir -p gpurun_out
for st in multi bench; do
  timeout -s KILL 180 python scripts/tc_debug.py $st > gpurun_out/tc_$st.log 2>&1
  echo "stage $st exit $?" >> gpurun_out/tc_$st.log
  tail -22 gpurun_out/tc_$st.log
done
for st in vit bench; do
  timeout -s KILL 120 python scripts/attn_debug.py $st > gpurun_out/attn_$st.log 2>&1
  echo "stage $st exit $?" >> gpurun_out/attn_$st.log
  tail -12 gpurun_out/attn_$st.log
done
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_bench.log
