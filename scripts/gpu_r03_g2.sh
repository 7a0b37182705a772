#!/bin/bash
# round 2, session 2: where the end-to-end loop loses its ~4 % against resident inputs (scripts/e2e_gap_probe.py)
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/e2e_gap_probe.py 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/r03g2_e2e_gap.log
