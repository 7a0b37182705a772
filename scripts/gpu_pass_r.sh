#!/bin/bash
# pass R: programmatic dependent launch
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/r_ops.log 2>&1; echo "ops rc=$?"
tail -3 gpurun_out/r_ops.log
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/r_e2e.log 2>&1; echo "e2e rc=$?"
tail -3 gpurun_out/r_e2e.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r_bench.log | cut -c1-330
EDGECAPE_PDL=0 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_nopdl.log 2>&1; echo "bench(no pdl) rc=$?"
tail -1 gpurun_out/r_bench_nopdl.log | cut -c1-330
