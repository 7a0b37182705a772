#!/bin/bash
# round 2, session 2, evidence pass on the final kernels (F16F8 rows with interleaved e4m3 planes): launch list of two eager
# steps, ncu --set full of one ViT layer's kernels inside the step (GEMMs, attention, LayerNorm)
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1000 -c 600 --csv --log-file gpurun_out/r03v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --sustained-seconds 0 > gpurun_out/r03v_ncu_bench.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/r03v_launches.csv > gpurun_out/r03v_launches.md 2>&1; head -12 gpurun_out/r03v_launches.md
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_f16x3|attention_tc_ps|layernorm_vec_kernel<6>" -s 40 -c 7 -o gpurun_out/prof_r03v_vit_layer python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --sustained-seconds 0 > gpurun_out/r03v_ncu_layer.log 2>&1; echo "ncu layer rc=$?"
ls -la gpurun_out/*.ncu-rep
