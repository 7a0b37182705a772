#!/bin/bash
# pass O: elect.sync issue of TMA / MMA, P-in-TMEM attention with wide P V MMAs
mkdir -p gpurun_out
timeout -s KILL 60 scripts/probes/_bin/umma_probe > gpurun_out/umma_probe.log 2>&1; grep -E "elect" gpurun_out/umma_probe.log
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/o_ops.log 2>&1; echo "ops rc=$?"
tail -3 gpurun_out/o_ops.log
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/o_e2e.log 2>&1; echo "e2e rc=$?"
tail -3 gpurun_out/o_e2e.log
timeout -s KILL 240 python scripts/attn_debug.py trace > gpurun_out/o_trace.log 2>&1; echo "trace rc=$?"
head -11 gpurun_out/o_trace.log
timeout -s KILL 240 python scripts/attn_debug.py tma 2>&1 | grep bench > gpurun_out/o_abench.log
timeout -s KILL 240 python scripts/attn_debug.py bench >> gpurun_out/o_abench.log 2>&1
cat gpurun_out/o_abench.log
timeout -s KILL 240 python scripts/tc_debug.py epi > gpurun_out/o_epi.log 2>&1; echo "epi rc=$?"
cat gpurun_out/o_epi.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/o_bench.log | cut -c1-330
