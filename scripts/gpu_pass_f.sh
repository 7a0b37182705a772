#!/bin/bash
# pass F: 512-thread attention, TC GCN (+bench), skeleton side stream, warm ncu launch list
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/attn_debug.py vit 2>&1 | tee gpurun_out/attn_vit.log | cut -c1-80
timeout -s KILL 120 python scripts/attn_debug.py bench 2>&1 | tee gpurun_out/attn_bench.log
timeout -s KILL 120 python scripts/gcn_bench.py 2>&1 | tee gpurun_out/gcn_bench.log
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1500 -c 450 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_bench.log
