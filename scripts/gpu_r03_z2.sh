#!/bin/bash
# round 2, session 2, final validation after the e4m3-plane interleave: compute-sanitizer memcheck over every F16F8 producer /
# consumer (scripts/gemm_f8_check.py), then pass Z (full GPU suite, smoke, bench with the driver's flags, reference arm)
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/gemm_f8_check.py 2>&1 | tail -2
timeout -s KILL 600 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python scripts/gemm_f8_check.py > gpurun_out/r03z2_memcheck_gemm_f8.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r03z2_memcheck_gemm_f8.log | cut -c1-200
sed -e 's/r03z_/r03z2_/g' scripts/gpu_r03_z.sh > /tmp/z2.sh; bash /tmp/z2.sh
