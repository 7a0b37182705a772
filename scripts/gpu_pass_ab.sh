#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "gcn" -x > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/ab_pytest.log
timeout -s KILL 200 python scripts/gcn_trace.py > gpurun_out/ab_gcn_trace.log 2>&1; echo "trace rc=$?"; head -34 gpurun_out/ab_gcn_trace.log
timeout -s KILL 200 python scripts/gcn_bench.py 2>&1 | head -5
