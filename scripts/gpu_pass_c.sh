#!/bin/bash
# pass C: coalesced TC epilogue + fused split producers + CUDA graphs
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/tc_debug.py multi > gpurun_out/tc_multi.log 2>&1; tail -3 gpurun_out/tc_multi.log
timeout -s KILL 120 python scripts/tc_debug.py bench > gpurun_out/tc_bench.log 2>&1; tail -7 gpurun_out/tc_bench.log
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1
tail -2 gpurun_out/bench_nograph.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_bench.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 40 -c 3 -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/ncu_full.log; tail -2 gpurun_out/ncu_full.log
