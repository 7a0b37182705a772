#!/bin/bash
# round 2 pass Q: ncu --set full (source-level stalls) of the fc1 GEMM with the TMA-store epilogue
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 30 -c 1 -o gpurun_out/prof_r02q_fc1 python scripts/gemm_f8_probe.py 0 1 > gpurun_out/r02q_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02q_ncu.log
