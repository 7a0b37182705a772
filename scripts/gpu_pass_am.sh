#!/bin/bash
# pass AM: per-GPU slices of configs[3] (5-shot, 8 queries / GPU) and configs[4] (ViT-L/14 @384, K=200, 4 queries / GPU)
mkdir -p gpurun_out
timeout -s KILL 150 python bench.py --shots 5 --batch 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/am_c4.log 2>&1; echo "c4 rc=$?"
tail -1 gpurun_out/am_c4.log | cut -c1-400
timeout -s KILL 150 python bench.py --backbone dinov2_vitl14 --image-size 384 --kpts 200 --batch 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/am_c5.log 2>&1; echo "c5 rc=$?"
tail -1 gpurun_out/am_c5.log | cut -c1-400
