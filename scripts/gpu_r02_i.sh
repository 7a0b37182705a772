#!/bin/bash
# round 2 pass I: persistent pipelined attention kernel (ViT blocks)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" > gpurun_out/r02i_pytest_attn.log 2>&1; echo "attention pytest rc=$?"; tail -12 gpurun_out/r02i_pytest_attn.log
timeout -s KILL 120 python scripts/attn_debug.py trace > gpurun_out/r02i_attn_trace.log 2>&1; head -14 gpurun_out/r02i_attn_trace.log
timeout -s KILL 120 python scripts/attn_debug.py tmabench > gpurun_out/r02i_attn_bench.log 2>&1; cat gpurun_out/r02i_attn_bench.log
timeout -s KILL 120 python scripts/attn_debug.py tmabench@3 >> gpurun_out/r02i_attn_bench.log 2>&1; tail -1 gpurun_out/r02i_attn_bench.log
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02i_pytest.log
timeout -s KILL 500 python bench.py --sustained-seconds 3 --no-cpu-baseline > gpurun_out/r02i_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r02i_bench.log > gpurun_out/r02i_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02i_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('nsk', d.get('north_star_kernels',{}).get('vit_attention'))
    print('sustained', d.get('sustained'))
except Exception as e:
    print('bench parse failed', e)
PY
