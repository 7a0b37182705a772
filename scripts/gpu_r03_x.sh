#!/bin/bash
# round 2, session 2, pass X: (a) the backbone as two independent chains (query / support images) on two streams,
# (b) the proj GEMM on the e4m3 kernel again now that its planes arrive in 128-byte-row boxes -- same box, alternating
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 2>/dev/null | tail -1 > gpurun_out/r03x_bench_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03x_bench_$name.json'))
    print('$name', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'], d['clocks']['sm_mhz'])
except Exception as e:
    print('$name bench parse failed', e)
PY
}
run default A=1
run chains2 EDGECAPE_VIT_CHAINS=2
run projf8 EDGECAPE_PROJ_F8=1
run default_b A=1
run chains2_b EDGECAPE_VIT_CHAINS=2
run chains2_projf8 EDGECAPE_VIT_CHAINS=2 EDGECAPE_PROJ_F8=1
EDGECAPE_VIT_CHAINS=2 timeout -s KILL 300 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x 2>&1 | tail -2
