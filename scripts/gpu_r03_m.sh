#!/bin/bash
# (and the GCN op tests on the restructured set-up barrier)
# round 2, session 2, pass M: compute-sanitizer on the project-first fused GCN (one CTA per item, clusters of two, persistent)
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "gcn" > gpurun_out/r03m_pytest_gcn.log 2>&1; echo "pytest gcn rc=$?"; tail -2 gpurun_out/r03m_pytest_gcn.log
timeout -s KILL 60 python scripts/gcn2_check.py > gpurun_out/r03m_check.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r03m_check.log
timeout -s KILL 600 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python scripts/gcn2_check.py --big > gpurun_out/r03m_memcheck_gcn2.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r03m_memcheck_gcn2.log
timeout -s KILL 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/gcn2_check.py > gpurun_out/r03m_racecheck_gcn2.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/r03m_racecheck_gcn2.log | cut -c1-200
timeout -s KILL 400 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/gcn2_check.py > gpurun_out/r03m_synccheck_gcn2.log 2>&1; echo "synccheck rc=$?"; tail -5 gpurun_out/r03m_synccheck_gcn2.log | cut -c1-200
