#!/bin/bash
# round 2, session 2, pass I: evidence for profiles/ -- full GPU suite, smoke, launch list of one eager step, ncu --set full
# of the project-first fused GCN, default bench (CPU baseline + parity + sustained)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r03i_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r03i_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1000 -c 600 --csv --log-file gpurun_out/r03i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --sustained-seconds 0 > gpurun_out/r03i_ncu_bench.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/r03i_launches.csv > gpurun_out/r03i_launches.md 2>&1; head -14 gpurun_out/r03i_launches.md
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gcn_fused2 -s 2 -c 1 -o gpurun_out/prof_r03i_gcn2 python scripts/gcn_once.py > gpurun_out/r03i_ncu_gcn.log 2>&1; echo "ncu gcn rc=$?"
timeout -s KILL 600 python bench.py > gpurun_out/r03i_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r03i_bench.log > gpurun_out/r03i_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03i_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','frac_of_sustained_peak','kernel_ms_per_step')})
    print('parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'], 'cpu', d['cpu_baseline']['value'])
    print('sustained', d['sustained']['value'], d['sustained']['clocks'])
    print('nsk', {k:(v['us'],v['frac']) for k,v in d['north_star_kernels'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
