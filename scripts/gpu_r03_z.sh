#!/bin/bash
# round 2, session 2, final validation: full GPU suite, smoke, default bench on the last commit
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r03z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r03z_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r03z_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r03z_bench.log > gpurun_out/r03z_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03z_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','frac_of_sustained_peak','kernel_ms_per_step')})
    print('parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'], 'cpu', d['cpu_baseline']['value'])
    print('sustained', d['sustained']['value'], d['sustained']['clocks'])
    print('nsk', {k:(v['us'],v['frac']) for k,v in d['north_star_kernels'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
