#!/bin/bash
# pass AI: launch list of the step with the one-kernel GCN (two eager steps' worth of launches)
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1000 -c 664 --csv --log-file gpurun_out/ai_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ai_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/ai_launches.csv | head -14
