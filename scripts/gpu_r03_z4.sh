#!/bin/bash
# round 2, session 2, last pass: MMA-thread wait trace of the large GEMMs on the final operand format, proj tile modes, then
# the bench with the driver's flags (new key f16f8_range_events) and smoke
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 2>&1 | head -8 | tee gpurun_out/r03z4_gemm_probe.log
timeout -s KILL 200 python scripts/proj_tile_probe.py 2>&1 | tail -8 | tee gpurun_out/r03z4_proj_tiles.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r03z4_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r03z4_bench.log > gpurun_out/r03z4_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03z4_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'], 'cpu', d['cpu_baseline']['value'], 'range events', d['f16f8_range_events'])
    print('sustained', d['sustained']['value'], 'roofline frac', d['roofline']['frac'])
except Exception as e:
    print('bench parse failed', e)
PY
