#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/gcn_bench.py 2>&1 | tee gpurun_out/gcn_bench.log | head -4
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log | cut -c1-1200
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 30 -c 4 -o gpurun_out/prof_j python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/ncu_full.log; tail -2 gpurun_out/ncu_full.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1400 -c 420 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
