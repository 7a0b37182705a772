#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gcn_fused -s 3 -c 2 -f -o gpurun_out/prof_ad_gcn_fused python scripts/gcn_once.py > gpurun_out/ad_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ad_ncu.log
