#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/e2e_probe.py > gpurun_out/r02m_e2e_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/r02m_e2e_probe.log | tail -6
