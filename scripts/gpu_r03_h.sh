#!/bin/bash
# round 2, session 2, pass H: attention writes F16F8 rows, the ViT proj GEMM on ec_gemm_f16f8 -- op tests, goldens, bench A/B
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" > gpurun_out/r03h_pytest_attn.log 2>&1; echo "pytest attention rc=$?"; tail -3 gpurun_out/r03h_pytest_attn.log
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x > gpurun_out/r03h_pytest_e2e.log 2>&1; echo "pytest e2e rc=$?"; tail -3 gpurun_out/r03h_pytest_e2e.log
for f in 1 0; do
EDGECAPE_PROJ_F8=$f timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/r03h_bench_proj_f8_$f.log 2>&1; echo "bench proj_f8=$f rc=$?"
tail -1 gpurun_out/r03h_bench_proj_f8_$f.log > gpurun_out/r03h_bench_proj_f8_$f.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03h_bench_proj_f8_$f.json'))
    print('proj_f8=$f', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'], 'parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'])
    print('  roofline', {k: d['roofline'][k] for k in ('achieved','frac','frac_of_sustained_peak','kernel_ms_per_step')})
    print('  nsk', {k:(round(v['us'],2),round(v['frac'],4)) for k,v in d['north_star_kernels'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
done
