#!/bin/bash
# round 2 pass C: full suite with the F16F8 default + f8-everywhere goldens, GEMM experiment flags, launch list
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02c_pytest.log
timeout -s KILL 300 python scripts/gemm_f8_probe.py > gpurun_out/r02c_f8_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02c_f8_probe.log | tail -8
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1000 -c 664 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --sustained-seconds 0 > gpurun_out/r02c_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/r02c_launches.csv > gpurun_out/r02c_launches.md 2>&1; head -20 gpurun_out/r02c_launches.md
