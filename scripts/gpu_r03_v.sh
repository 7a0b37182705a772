#!/bin/bash
# round 2, session 2, pass V: same-box A/B of the GEMM before / after the K-split code paths were added (default: off)
mkdir -p gpurun_out
cp edgecape_b200/libedgecape_b200.so /tmp/cur.so
for round in 1 2; do
  cp edgecape_b200/libedgecape_b200_prev.so edgecape_b200/libedgecape_b200.so
  echo "== before (round $round)"; timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 2>&1 | grep -v "MMA thread" | head -3
  cp /tmp/cur.so edgecape_b200/libedgecape_b200.so
  echo "== after (round $round)"; timeout -s KILL 200 python scripts/gemm_f8_probe.py 0 2>&1 | grep -v "MMA thread" | head -3
done
