#!/bin/bash
# pass AL: fill loads hoisted above the setup barrier -- GCN tests, end-to-end goldens, phase trace, micro-benchmark
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py tests/test_e2e_gpu.py -q -m gpu -k "gcn or e2e or golden or detector" > gpurun_out/al_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/al_pytest.log
timeout -s KILL 200 python scripts/gcn_trace.py > gpurun_out/al_gcn_trace.log 2>&1; echo "trace rc=$?"; head -9 gpurun_out/al_gcn_trace.log; grep -E "ACC ready|w:end" gpurun_out/al_gcn_trace.log | head -2
timeout -s KILL 200 python scripts/gcn_bench.py 2>&1 | head -5 | tee gpurun_out/al_gcn_bench.jsonl | cut -c1-190
