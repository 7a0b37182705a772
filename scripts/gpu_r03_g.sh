#!/bin/bash
# round 2, session 2, pass G: where does the GEMM's MMA thread wait (qkv / fc1 / fc2 at the ViT-B bench shapes)?
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 > gpurun_out/r03g_gemm_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r03g_gemm_probe.log
