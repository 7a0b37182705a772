#!/bin/bash
for mode in head all none; do
  echo -n "pdl=$mode: "
  EC_PDL_MODE=$mode timeout -s KILL 200 python scripts/overlap_probe.py 2>&1 | tail -1
done
timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/w_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/w_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'eager', round(d['eager_ms_per_step'],2), d['clocks']['reasons'])"
