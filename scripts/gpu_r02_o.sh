#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 60 scripts/probes/_bin/l2_feed_probe > gpurun_out/r02o_l2_feed_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/r02o_l2_feed_probe.log
timeout -s KILL 200 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "markov or edge_weights" 2>&1 | tail -3
