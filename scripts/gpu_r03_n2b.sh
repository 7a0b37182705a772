#!/bin/bash
# round 2, session 2: configs[1] on 2 GPUs under the driver's torchrun line, final code (interleaved e4m3 planes)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -2
timeout -s KILL 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03n2b_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r03n2b_bench.log > gpurun_out/r03n2b_bench.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03n2b_bench.json'))
    print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['max_rel_err'], d['parity']['argmax_equal'], d['clocks'])
except Exception as e:
    print('bench parse failed', e)
PY
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-300
