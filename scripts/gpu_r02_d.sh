#!/bin/bash
# round 2 pass D: F16F8 pair-GEMM experiment flags (one process per flag), ncu --set full of one ViT layer's GEMMs inside the
# step, the config presets c4 / c5 on one GPU, memcheck on the F16F8 tests
mkdir -p gpurun_out
for f in 0 4 8 1 9; do
  timeout -s KILL 60 python scripts/gemm_f8_probe.py $f > gpurun_out/r02d_f8_probe_$f.log 2>&1; echo "probe flag $f rc=$?"; cat gpurun_out/r02d_f8_probe_$f.log | tail -6
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 25 -c 4 -o gpurun_out/prof_r02d_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --sustained-seconds 0 > gpurun_out/r02d_ncu.log 2>&1; echo "ncu rc=$?"
for c in c4 c5; do
  timeout -s KILL 600 python bench.py --config $c --sustained-seconds 3 > gpurun_out/r02d_bench_$c.log 2>&1; echo "bench $c rc=$?"
  tail -1 gpurun_out/r02d_bench_$c.log > gpurun_out/r02d_bench_$c.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02d_bench_$c.json'))
    print('$c', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config']['workload'][:60], d['gpu_launches'])
    print('   parity', d.get('parity', {}).get('max_rel_err'), d.get('parity', {}).get('argmax_equal'), 'cpu', d.get('cpu_baseline', {}).get('value'))
except Exception as e:
    print('bench parse failed', e)
PY
done
timeout -s KILL 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_f8.py -q -m gpu -x -k "epilogues or range or planes or (matches_fp64 and 1300)" > gpurun_out/r02d_memcheck_f8.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02d_memcheck_f8.log
