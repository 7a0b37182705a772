#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 180 python scripts/attn_debug.py tma 2>&1 | tee gpurun_out/attn_tma.log | cut -c1-100
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_f16x3_kernel<256, 1>|attention_tc_tma_kernel" -s 10 -c 5 -o gpurun_out/prof_i python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/ncu_full.log; tail -2 gpurun_out/ncu_full.log
