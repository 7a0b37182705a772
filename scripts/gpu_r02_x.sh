#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "hop_bias or attention" > gpurun_out/r02x_pytest_attn.log 2>&1; echo "attn pytest rc=$?"; tail -3 gpurun_out/r02x_pytest_attn.log
EDGECAPE_HOP_FUSED=1 timeout -s KILL 600 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x -k "tcgen05 and not f8 and not simt" > gpurun_out/r02x_pytest_hopfused.log 2>&1; echo "e2e with hop fused rc=$?"; tail -3 gpurun_out/r02x_pytest_hopfused.log
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02x_pytest.log
