#!/bin/bash
# round 2, session 2, pass B: phase trace experiments of the project-first fused GCN
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/gcn2_trace.py --experiments --short > gpurun_out/r03b_gcn2_trace.log 2>&1; echo "trace rc=$?"; grep -E "^B=|stored|GEMM A issued|epilogue|arrived" gpurun_out/r03b_gcn2_trace.log
