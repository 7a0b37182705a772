#!/bin/bash
# pass M: warp-specialised attention kernels (control warp issues TMA/MMA)
mkdir -p gpurun_out
timeout -s KILL 240 python scripts/attn_debug.py tma > gpurun_out/m_attn.log 2>&1; echo "attn rc=$?"
tail -14 gpurun_out/m_attn.log
timeout -s KILL 240 python scripts/attn_debug.py trace > gpurun_out/m_trace.log 2>&1; echo "trace rc=$?"
head -12 gpurun_out/m_trace.log
timeout -s KILL 240 python scripts/attn_debug.py vit > gpurun_out/m_vit.log 2>&1; echo "vit rc=$?"
tail -5 gpurun_out/m_vit.log
timeout -s KILL 240 python scripts/attn_debug.py bench > gpurun_out/m_abench.log 2>&1; echo "abench rc=$?"
tail -6 gpurun_out/m_abench.log
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" > gpurun_out/m_ops.log 2>&1; echo "ops rc=$?"
tail -3 gpurun_out/m_ops.log
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/m_e2e.log 2>&1; echo "e2e rc=$?"
tail -3 gpurun_out/m_e2e.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/m_bench.log | cut -c1-400
