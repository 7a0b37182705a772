#!/bin/bash
# round 2 pass A: new parity pins (c2-b16 pipelined, c4 ViT-B 5-shot), smoke at ViT-S/224, default bench with parity /
# sustained keys, first hardware run of the fp16+e4m3 GEMM prototype, memcheck on the tcgen05 kernels
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest.log
EDGECAPE_TEST_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_experimental_gpu.py -q -m gpu > gpurun_out/r02a_pytest_exp.log 2>&1; echo "exp rc=$?"; tail -15 gpurun_out/r02a_pytest_exp.log
timeout -s KILL 200 python scripts/gemm_f8x_bench.py > gpurun_out/r02a_f8x_bench.log 2>&1; echo "f8x rc=$?"; cat gpurun_out/r02a_f8x_bench.log | tail -8
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 500 python bench.py > gpurun_out/r02a_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r02a_bench.log > gpurun_out/r02a_bench.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02a_bench.json'))
    print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['gpu_launches'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','frac_of_sustained_peak')})
    print('parity', d.get('parity'))
    print('sustained', d.get('sustained'))
    print('nsk', d.get('north_star_kernels'))
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout -s KILL 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "test_gemm_tc_fp32_grade or test_attention_tma_on_split_qkv or test_gcn_fused_kernel" > gpurun_out/r02a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02a_memcheck.log
