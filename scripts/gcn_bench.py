"""GCN feed-forward micro-benchmark (north-star target: fraction of the HBM roofline at batch 64)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from edgecape_b200 import ops  # noqa: E402


def run(B, K, d, dff, tc, iters=50, fused=False):
    D = torch.device("cuda")
    ops.TENSOR_CORES = tc
    ops.GCN_FUSED = fused
    if fused and not ops.gcn_fused_ok(B, K, d, dff):
        return
    x = torch.randn(B, K, d, device=D)
    U = torch.rand(B, K, K, device=D)
    mask = torch.zeros(B, K, dtype=torch.uint8, device=D)
    adj = ops.soft_normalize_adj(U, mask)
    W = torch.randn(2 * dff, d, device=D) * d ** -0.5
    bias = torch.randn(2 * dff, device=D) * 0.1
    Wp = ops.gcn_pack_weights(W, bias)
    out = torch.empty(B, K, dff, device=D)
    for _ in range(5):
        ops.gcn(x, adj, Wp, out=out)
    g = torch.cuda.CUDAGraph()            # graph replay: kernel time, not Python launch time
    with torch.cuda.graph(g):
        for _ in range(iters):
            ops.gcn(x, adj, Wp, out=out)
    g.replay()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / iters
    bytes_alg = B * K * d * 4 + B * K * K * 4 + B * K + (2 * dff * d + 2 * dff) * 4 + B * K * dff * 4
    flops = 2.0 * B * K * d * 2 * dff + 2.0 * B * K * K * d
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
    gbs = bytes_alg / us / 1e3
    print(json.dumps(dict(op="gcn", B=B, K=K, d=d, dff=dff, path=({2: "fused2_project_first", 1: "fused_aggregate_first", 0: "tcgen05_pair"}[int(fused)]) if tc else "simt_fp32", us=round(us, 2),
                          algorithmic_MB=round(bytes_alg / 1e6, 2), achieved_GBs=round(gbs, 1), hbm_peak_GBs=peak,
                          frac_hbm=round(gbs / peak, 4), algorithmic_TFLOPs=round(flops / us / 1e6, 1))), flush=True)


if __name__ == "__main__":
    for fused in (2, 1):
        for B in (64, 16, 74, 148, 2048):
            run(B, 100, 256, 384, True, iters=50 if B < 1000 else 5, fused=fused)
        run(64, 100, 256, 768, True, fused=fused)
    if "--k200" in sys.argv:      # configs[4]: 200 keypoints -- clusters of two CTAs vs the two-launch path
        for B in (4, 32, 256):
            for fused in (2, 0):
                run(B, 200, 256, 384, True, iters=20, fused=fused)
        for fused in (2, 0):
            run(4, 200, 256, 1024, True, iters=20, fused=fused)
    if "--fused-only" in sys.argv:
        sys.exit(0)
    for tc in (True, False):
        run(64, 100, 256, 384, tc)
        run(64, 100, 256, 768, tc)
        run(32, 200, 256, 384, tc)
        run(2048, 100, 256, 384, tc, iters=5)
