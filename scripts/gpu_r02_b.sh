#!/bin/bash
# round 2 pass B: the F16F8 GEMM integrated into the main kernel (all tile modes incl. the CTA pair), on the ViT by default
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gemm_f8.py -q -m gpu -x > gpurun_out/r02b_pytest_f8.log 2>&1; echo "f8 pytest rc=$?"; tail -15 gpurun_out/r02b_pytest_f8.log
timeout -s KILL 200 python scripts/gemm_f8x_bench.py > gpurun_out/r02b_f8x_bench.log 2>&1; echo "f8x rc=$?"; cat gpurun_out/r02b_f8x_bench.log | tail -8
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest.log
for f8 in 1 0; do
EDGECAPE_GEMM_F8=$f8 timeout -s KILL 500 python bench.py --sustained-seconds 3 --no-cpu-baseline > gpurun_out/r02b_bench_f8_$f8.log 2>&1; echo "bench f8=$f8 rc=$?"
tail -1 gpurun_out/r02b_bench_f8_$f8.log > gpurun_out/r02b_bench_f8_$f8.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02b_bench_f8_$f8.json'))
    print('f8=$f8', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','frac_of_sustained_peak','kernel_ms_per_step')})
    print('sustained', d.get('sustained'))
except Exception as e:
    print('bench parse failed', e)
PY
done
