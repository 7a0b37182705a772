#!/bin/bash
# round 2, session 2, pass S: X ahead of W in the first loads of the fused GCN, end-of-kernel store waits trimmed to their reads -- op tests, standalone timings, bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_f8.py -q -m gpu -x > gpurun_out/r03s_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -2 gpurun_out/r03s_pytest_ops.log
timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 > gpurun_out/r03s_gemm_probe.log 2>&1; echo "probe rc=$?"; grep -v "MMA thread" gpurun_out/r03s_gemm_probe.log
timeout -s KILL 300 python scripts/gcn_bench.py --fused-only 2>&1 | head -6 | cut -c1-150
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/r03s_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r03s_bench.log > gpurun_out/r03s_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03s_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'], 'parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','kernel_ms_per_step')})
    print('nsk', {k:(round(v['us'],2),round(v['frac'],4)) for k,v in d['north_star_kernels'].items()})
except Exception as e:
    print('bench parse failed', e)
PY
