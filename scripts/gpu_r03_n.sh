#!/bin/bash
# round 2, session 2: the driver's torchrun line at N GPUs for configs[1] (default), configs[3] (--config c4) and configs[4] (--config c5)
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt
p=29521
for c in c2 c4 c5; do
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p bench.py --gpus $N --steps 10 --warmup 3 --config $c --sustained-seconds 0 > gpurun_out/r03_n${N}_$c.log 2>&1
  echo "$c exit $?"
  grep -E "^\{" gpurun_out/r03_n${N}_$c.log | tail -1 > gpurun_out/r03_n${N}_$c.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03_n${N}_$c.json'))
    print('$c', 'n_gpus', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config']['workload'][:50])
except Exception as e:
    print('parse failed', e)
PY
  p=$((p+1))
done
