#!/bin/bash
# round 2, session 2, pass Q: GEMM with a static first tile (dynamic from the second) -- tests, standalone timings, bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_f8.py -q -m gpu -x -k "gemm or split_only or linear" > gpurun_out/r03q_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -2 gpurun_out/r03q_pytest_gemm.log
timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 > gpurun_out/r03q_gemm_probe.log 2>&1; echo "probe rc=$?"; grep -v "MMA thread" gpurun_out/r03q_gemm_probe.log
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x -k "benchmarked or c2_vitb" > gpurun_out/r03q_pytest_e2e.log 2>&1; echo "pytest e2e rc=$?"; tail -2 gpurun_out/r03q_pytest_e2e.log
