#!/bin/bash
# round 2, session 2, pass K: fused GCN for K > 128 (clusters of two CTAs, T1 rows exchanged through distributed shared memory)
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "gcn" > gpurun_out/r03k_pytest_gcn.log 2>&1; echo "pytest gcn rc=$?"; tail -12 gpurun_out/r03k_pytest_gcn.log | cut -c1-220
