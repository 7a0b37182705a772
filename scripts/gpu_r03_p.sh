#!/bin/bash
# round 2, session 2, pass P: attention (persistent kernel) with the softmax's TMEM reads one chunk ahead; tests + trace + timing
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_f8.py -q -m gpu -x -k "attention or interleaved" > gpurun_out/r03p_pytest_attn.log 2>&1; echo "pytest attention rc=$?"; tail -2 gpurun_out/r03p_pytest_attn.log
for i in 1 2 3; do timeout -s KILL 100 python scripts/attn_debug.py tmabench 2>&1 | tail -1; done
timeout -s KILL 100 python scripts/attn_debug.py trace > gpurun_out/r03p_attention_trace.log 2>&1; head -12 gpurun_out/r03p_attention_trace.log
