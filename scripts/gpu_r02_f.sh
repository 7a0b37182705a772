#!/bin/bash
# round 2 pass F: decompose the GEMM epilogue cost (experiment flags 16 = no split stores, 32 = no GELU)
mkdir -p gpurun_out
for f in 0 16 32 48 8; do
  timeout -s KILL 60 python scripts/gemm_f8_probe.py $f > gpurun_out/r02f_f8_probe_$f.log 2>&1; echo "probe flag $f rc=$?"; cat gpurun_out/r02f_f8_probe_$f.log | head -3
done
