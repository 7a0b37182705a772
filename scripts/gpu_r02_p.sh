#!/bin/bash
# round 2 pass P: split-only epilogue with direct register stores (no shared-memory traffic) against the TMA-store form
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gemm_f8.py -q -m gpu -x -k "tma_stores" > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p_pytest.log
for m in 1 2 0; do
  timeout -s KILL 60 python scripts/gemm_f8_probe.py 0 $m > gpurun_out/r02p_probe_mode$m.log 2>&1; echo "mode $m rc=$?"; head -3 gpurun_out/r02p_probe_mode$m.log
done
