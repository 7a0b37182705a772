#!/bin/bash
# round 2 pass H: split-only GEMM epilogue through TMA stores
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gemm_f8.py tests/test_ops_gpu.py -q -m gpu -x -k "gemm or linear or fused_split or tma_stores" > gpurun_out/r02h_pytest_gemm.log 2>&1; echo "gemm pytest rc=$?"; tail -12 gpurun_out/r02h_pytest_gemm.log
for f in 0 16 8; do
  timeout -s KILL 60 python scripts/gemm_f8_probe.py $f > gpurun_out/r02h_f8_probe_$f.log 2>&1; echo "probe flag $f rc=$?"; cat gpurun_out/r02h_f8_probe_$f.log | head -6
done
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02h_pytest.log
timeout -s KILL 500 python bench.py --sustained-seconds 3 --no-cpu-baseline > gpurun_out/r02h_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r02h_bench.log > gpurun_out/r02h_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02h_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','frac_of_sustained_peak','kernel_ms_per_step')})
    print('sustained', d.get('sustained'))
except Exception as e:
    print('bench parse failed', e)
PY
