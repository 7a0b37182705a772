#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "gcn" -x > gpurun_out/ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ac_pytest.log
timeout -s KILL 200 python scripts/gcn_trace.py > gpurun_out/ac_gcn_trace.log 2>&1; echo "trace rc=$?"; head -8 gpurun_out/ac_gcn_trace.log; grep -E "ACC ready|w:end" gpurun_out/ac_gcn_trace.log | head -2
timeout -s KILL 200 python scripts/gcn_bench.py 2>&1 | head -5 | tee gpurun_out/ac_gcn_bench.jsonl
EDGECAPE_GCN_FUSED=1 timeout -s KILL 600 python -m pytest tests/test_e2e_gpu.py -q -m gpu -x > gpurun_out/ac_e2e_fused.log 2>&1; echo "e2e fused rc=$?"; tail -3 gpurun_out/ac_e2e_fused.log
for f in 0 1; do
EDGECAPE_GCN_FUSED=$f timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/ac_bench_fused$f.log 2>&1; echo "bench fused=$f rc=$?"
tail -1 gpurun_out/ac_bench_fused$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['reasons'], d['gpu_launches'])"
done
