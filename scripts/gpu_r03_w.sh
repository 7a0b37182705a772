#!/bin/bash
# round 2, session 2, pass W: would 128-byte-row TMA boxes for the e4m3 planes help the GEMM?  (timing experiment, dbg 64)
mkdir -p gpurun_out
for f in 0 64 0 64; do echo "== dbg $f"; timeout -s KILL 200 python scripts/gemm_f8_probe.py $f 2>&1 | grep -v "MMA thread" | head -3; done
