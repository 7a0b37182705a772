#!/bin/bash
# pass Q: head attentions through the split-fp16 TMA / P-in-TMEM kernel (head dim 32, key mask, bias)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/q_ops.log 2>&1; echo "ops rc=$?"
tail -12 gpurun_out/q_ops.log
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/q_e2e.log 2>&1; echo "e2e rc=$?"
tail -8 gpurun_out/q_e2e.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/q_bench.log | cut -c1-330
