#!/bin/bash
# pass S: evidence -- full bench line, GCN bench, ncu --set full of the P-in-TMEM attention and the CTA-pair GEMM
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/s_bench.log | cut -c1-300
timeout -s KILL 120 python scripts/gcn_bench.py > gpurun_out/s_gcn.log 2>&1; cat gpurun_out/s_gcn.log | cut -c1-250
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_ts -s 4 -c 2 -o gpurun_out/prof_s_attn python scripts/attn_debug.py tmabench > gpurun_out/s_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 12 -c 4 -o gpurun_out/prof_s_gemm python scripts/tc_debug.py epi > gpurun_out/s_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
