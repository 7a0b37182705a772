#!/bin/bash
# round 2 pass S: Markov powers kernel (row-block form) timing + tests; overlap probe
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "markov or edge_weights" 2>&1 | tail -2
timeout -s KILL 120 python - <<'PY' 2>&1 | tail -4
import sys, torch
sys.path.insert(0, ".")
from edgecape_b200 import ops
D = torch.device("cuda")
B, K, H = 16, 100, 4
P = torch.rand(B, K, K, device=D); P = P / P.sum(-1, keepdim=True)
hops = torch.zeros(H + 1, B, K, K, device=D); hops[0] = torch.eye(K, device=D); hops[1] = P
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
def gemms():
    for h in range(2, H + 1):
        ops.gemm(hops[h // 2], hops[h - h // 2], out=hops[h], b_kmajor=False)
print(f"markov powers B={B} K={K}: one kernel {t(lambda: ops.markov_powers_(hops)):.1f} us, three batched fp32 GEMMs {t(gemms):.1f} us")
PY
timeout -s KILL 300 python scripts/overlap_probe.py > gpurun_out/r02s_overlap.log 2>&1; tail -1 gpurun_out/r02s_overlap.log
