#!/bin/bash
# multi-GPU check: the exact launch line the driver uses
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_n$N.log
grep -E "^\{|exit|Error|error" gpurun_out/bench_n$N.log | cut -c1-900 | tail -5
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_ref_n$N.log
grep -E "^\{|exit" gpurun_out/bench_ref_n$N.log | cut -c1-300 | tail -3
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
grep -E "^\{" gpurun_out/bench_n1.log | cut -c1-200
