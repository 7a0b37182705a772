#!/bin/bash
# round 2 pass V: two P V issuing threads + earlier V loads in the persistent attention; mask fast path in the general kernel
mkdir -p gpurun_out
timeout -s KILL 100 python scripts/attn_debug.py trace32 > gpurun_out/r02v_attn_trace32.log 2>&1; grep -E "encoder-shaped|CTA total" gpurun_out/r02v_attn_trace32.log
timeout -s KILL 100 python scripts/attn_debug.py trace > gpurun_out/r02v_attn_trace.log 2>&1
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02v_pytest.log
timeout -s KILL 500 python bench.py --sustained-seconds 3 --no-cpu-baseline > gpurun_out/r02v_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r02v_bench.log > gpurun_out/r02v_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02v_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('nsk', {k:(round(v['us'],1),round(v['frac'],3)) for k,v in d['north_star_kernels'].items()}, d['north_star_kernels']['vit_attention']['issued_frac'])
    print('sustained', d.get('sustained'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout -s KILL 300 python scripts/overlap_probe.py > gpurun_out/r02v_overlap.log 2>&1; tail -1 gpurun_out/r02v_overlap.log
