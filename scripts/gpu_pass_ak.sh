#!/bin/bash
# pass AK: 2-GPU sanity of the driver's launch line after the bench / GCN changes
mkdir -p gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/ak_bench_n2.log 2>&1
echo "exit $?"
grep -E "^\{" gpurun_out/ak_bench_n2.log | tail -1 > gpurun_out/ak_bench_n2.json
python -c "import json; d=json.load(open('gpurun_out/ak_bench_n2.json')); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['reasons'], 'north_star' in str(d.keys()))"
