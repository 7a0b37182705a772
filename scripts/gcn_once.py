"""A few launches of the GCN feed-forward at the north-star shape (B=64, K=100, d=256, dff=384) for ncu captures."""
import sys

import torch

sys.path.insert(0, ".")
from edgecape_b200 import ops  # noqa: E402

fused = 0 if "--two-kernel" in sys.argv else (1 if "--aggregate-first" in sys.argv else 2)   # 2: gcn_fused2_tcgen05.cu
D = torch.device("cuda")
ops.TENSOR_CORES, ops.GCN_FUSED = True, fused
B, K, d, dff = 64, 100, 256, 384
x = torch.randn(B, K, d, device=D)
adj = ops.soft_normalize_adj(torch.rand(B, K, K, device=D), torch.zeros(B, K, dtype=torch.uint8, device=D))
Wp = ops.gcn_pack_weights(torch.randn(2 * dff, d, device=D) * d ** -0.5, torch.randn(2 * dff, device=D) * 0.1)
out = torch.empty(B, K, dff, device=D)
for _ in range(6):
    ops.gcn(x, adj, Wp, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
