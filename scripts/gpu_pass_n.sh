#!/bin/bash
# pass N: packed fp16 conversions, vector LayerNorm, residual prefetch in the GEMM epilogue
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/n_ops.log 2>&1; echo "ops rc=$?"
tail -3 gpurun_out/n_ops.log
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/n_e2e.log 2>&1; echo "e2e rc=$?"
tail -3 gpurun_out/n_e2e.log
timeout -s KILL 240 python scripts/attn_debug.py trace > gpurun_out/n_trace.log 2>&1; echo "trace rc=$?"
head -12 gpurun_out/n_trace.log
timeout -s KILL 240 python scripts/attn_debug.py bench > gpurun_out/n_abench.log 2>&1
timeout -s KILL 240 python scripts/attn_debug.py tma 2>&1 | grep bench >> gpurun_out/n_abench.log
cat gpurun_out/n_abench.log
timeout -s KILL 240 python scripts/tc_debug.py epi > gpurun_out/n_epi.log 2>&1; echo "epi rc=$?"
head -6 gpurun_out/n_epi.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/n_bench.log | cut -c1-330
