#!/bin/bash
# round 2, session 2: compute-sanitizer memcheck over every F16F8 producer / consumer after the e4m3-plane interleave
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/gemm_f8_check.py 2>&1 | tail -3
timeout -s KILL 600 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python scripts/gemm_f8_check.py > gpurun_out/r03z3_memcheck_gemm_f8.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r03z3_memcheck_gemm_f8.log | cut -c1-200
