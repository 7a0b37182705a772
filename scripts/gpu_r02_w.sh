#!/bin/bash
# round 2 pass W: hop-bias MLP fused into the decoder self-attention
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "hop_bias or attention" > gpurun_out/r02w_pytest_attn.log 2>&1; echo "attn pytest rc=$?"; tail -5 gpurun_out/r02w_pytest_attn.log
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02w_pytest.log
for hf in 1 0; do
EDGECAPE_HOP_FUSED=$hf timeout -s KILL 300 python scripts/overlap_probe.py > gpurun_out/r02w_overlap_$hf.log 2>&1; echo "hop_fused=$hf $(tail -1 gpurun_out/r02w_overlap_$hf.log)"
done
timeout -s KILL 500 python bench.py --sustained-seconds 3 --no-cpu-baseline > gpurun_out/r02w_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r02w_bench.log > gpurun_out/r02w_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02w_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
except Exception as e:
    print('bench parse failed', e)
PY
