#!/bin/bash
# pass AA: bring-up of the one-kernel GCN (gcn_fused_tcgen05.cu)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "gcn" -x > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/aa_pytest.log
timeout -s KILL 200 python scripts/gcn_bench.py > gpurun_out/aa_gcn_bench.jsonl 2> gpurun_out/aa_gcn_bench.err; echo "bench rc=$?"; cat gpurun_out/aa_gcn_bench.jsonl; tail -3 gpurun_out/aa_gcn_bench.err
