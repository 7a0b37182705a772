#!/bin/bash
# round 2 pass J: how much of the step is the head?  (backbone graph alone / head graph alone / both concurrently)
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/overlap_probe.py > gpurun_out/r02j_overlap.log 2>&1; echo "overlap rc=$?"; tail -8 gpurun_out/r02j_overlap.log
