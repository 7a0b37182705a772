#!/bin/bash
# pass X: final validation -- full GPU suite, ncu --set full of one ViT layer's four GEMMs inside the step, final bench
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x_pytest.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 25 -c 4 -o gpurun_out/prof_x_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/x_ncu.log 2>&1; echo "ncu rc=$?"
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/x_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/x_bench.log | cut -c1-250
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
