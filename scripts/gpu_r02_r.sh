#!/bin/bash
# round 2 pass R: bias / LayerScale prefetched (coalesced load + shuffles in the TMA epilogue), Markov powers kernel
mkdir -p gpurun_out
timeout -s KILL 60 python scripts/gemm_f8_probe.py 0 1 > gpurun_out/r02r_probe.log 2>&1; head -6 gpurun_out/r02r_probe.log
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02r_pytest.log
timeout -s KILL 500 python bench.py --sustained-seconds 3 --no-cpu-baseline > gpurun_out/r02r_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r02r_bench.log > gpurun_out/r02r_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02r_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('sustained', d.get('sustained'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout -s KILL 300 python scripts/overlap_probe.py > gpurun_out/r02r_overlap.log 2>&1; tail -1 gpurun_out/r02r_overlap.log
