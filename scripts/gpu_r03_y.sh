#!/bin/bash
# round 2, session 2, pass Y: F16F8 rows with the e4m3 planes interleaved per 64 columns (one 128-byte-row TMA box per operand
# and k-block) -- GEMM probe, then the full validation of pass Z on this tree
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 2>&1 | grep -v "MMA thread" | head -4 | tee gpurun_out/r03y_gemm_probe.log
sed -e 's/r03z_/r03y_/g' scripts/gpu_r03_z.sh > /tmp/y.sh; bash /tmp/y.sh
