#!/bin/bash
# round 2, session 2, pass T: K-split of the tail tiles in the GEMM (opt-in): tests, standalone timings on / off, bench on / off
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_f8.py tests/test_ops_gpu.py -q -m gpu -x -k "ksplit or gemm" > gpurun_out/r03t_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r03t_pytest.log | cut -c1-200
for k in 0 1; do
echo "== EDGECAPE_GEMM_KSPLIT=$k"
EDGECAPE_GEMM_KSPLIT=$k timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 2>&1 | grep -v "MMA thread" | head -8
done
