#!/bin/bash
# round 2, session 2: step anatomy on the final code (backbone graph alone, head graph alone, both concurrently)
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/overlap_probe.py 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/r03o_overlap_probe.log
