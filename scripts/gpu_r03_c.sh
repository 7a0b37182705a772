#!/bin/bash
# round 2, session 2, pass C: GCN tests + short trace + bench of the project-first fused GCN
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "gcn" > gpurun_out/r03c_pytest_gcn.log 2>&1; echo "pytest gcn rc=$?"; tail -4 gpurun_out/r03c_pytest_gcn.log
timeout -s KILL 200 python scripts/gcn2_trace.py --experiments > gpurun_out/r03c_gcn2_trace.log 2>&1; echo "trace rc=$?"; head -28 gpurun_out/r03c_gcn2_trace.log
timeout -s KILL 300 python scripts/gcn_bench.py --fused-only > gpurun_out/r03c_gcn_bench.jsonl 2>&1; echo "bench rc=$?"; cut -c1-200 gpurun_out/r03c_gcn_bench.jsonl
