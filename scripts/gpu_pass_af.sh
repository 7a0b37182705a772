#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/gcn_trace.py --experiments > gpurun_out/af_gcn_trace.log 2>&1; echo "trace rc=$?"; grep -E "^B=|^---|A1 tiles|X tiles|GEMM1 done|ACC ready|w:end" gpurun_out/af_gcn_trace.log
