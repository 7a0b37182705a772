#!/bin/bash
# pass P: two S-issuing warps + coalesced epilogue in the P-in-TMEM attention; GEMM tile modes per shape
mkdir -p gpurun_out
timeout -s KILL 240 python scripts/attn_debug.py tma > gpurun_out/p_attn.log 2>&1; echo "attn rc=$?"
tail -12 gpurun_out/p_attn.log | cut -c1-70
timeout -s KILL 240 python scripts/attn_debug.py trace > gpurun_out/p_trace.log 2>&1; echo "trace rc=$?"
head -11 gpurun_out/p_trace.log
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention or fused" > gpurun_out/p_ops.log 2>&1; echo "ops rc=$?"
tail -2 gpurun_out/p_ops.log
timeout -s KILL 300 python scripts/tc_debug.py modes > gpurun_out/p_modes.log 2>&1; echo "modes rc=$?"
cat gpurun_out/p_modes.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/p_bench.log | cut -c1-330
