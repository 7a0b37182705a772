#!/bin/bash
# pass L: two-key-block TMA attention (ViT-L/384), then parity + a c5-shaped bench
mkdir -p gpurun_out
timeout -s KILL 240 python scripts/attn_debug.py tma > gpurun_out/l_attn.log 2>&1; echo "attn rc=$?"
tail -15 gpurun_out/l_attn.log
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" > gpurun_out/l_ops.log 2>&1; echo "ops rc=$?"
tail -3 gpurun_out/l_ops.log
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py -x -q -m gpu > gpurun_out/l_e2e.log 2>&1; echo "e2e rc=$?"
tail -3 gpurun_out/l_e2e.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > gpurun_out/l_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/l_bench.log
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --backbone dinov2_vitl14 --image-size 384 --kpts 200 --batch 8 --no-cpu-baseline > gpurun_out/l_bench_c5.log 2>&1; echo "c5 rc=$?"
tail -1 gpurun_out/l_bench_c5.log
