#!/bin/bash
# round 2, session 2, pass A: first hardware run of the project-first fused GCN (gcn_fused2_tcgen05.cu)
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "gcn" > gpurun_out/r03a_pytest_gcn.log 2>&1; echo "pytest gcn rc=$?"; tail -15 gpurun_out/r03a_pytest_gcn.log
timeout -s KILL 200 python scripts/gcn2_trace.py --experiments > gpurun_out/r03a_gcn2_trace.log 2>&1; echo "trace rc=$?"; head -24 gpurun_out/r03a_gcn2_trace.log
timeout -s KILL 300 python scripts/gcn_bench.py --fused-only > gpurun_out/r03a_gcn_bench.jsonl 2>&1; echo "bench rc=$?"; cat gpurun_out/r03a_gcn_bench.jsonl
