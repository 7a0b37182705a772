"""One small launch of the project-first fused GCN per shape, checked against torch fp64 (for compute-sanitizer runs)."""
import sys

import torch

sys.path.insert(0, ".")
from edgecape_b200 import ops  # noqa: E402

D = torch.device("cuda")
ops.TENSOR_CORES, ops.GCN_FUSED = True, 2
for B, K, d, dff in ((3, 100, 256, 384), (2, 200, 256, 384), (150, 100, 256, 384) if "--big" in sys.argv else (2, 64, 64, 64)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, K, d, generator=g)
    adj = torch.rand(B, 2, K, K, generator=g)
    adj[:, 0] = torch.diag_embed(torch.rand(B, K, generator=g) + 0.5)
    W = torch.randn(2 * dff, d, generator=g) * d ** -0.5
    bias = torch.randn(2 * dff, generator=g) * 0.1
    Wp = ops.gcn_pack_weights(W.to(D), bias.to(D))
    for split in ("no", "only"):
        y = ops.gcn(x.to(D), adj.to(D).contiguous(), Wp, split=split)
        torch.cuda.synchronize()
        if split == "only":
            y = y.data[:, :dff].float() + y.data[:, y.Kp:y.Kp + dff].float()
        xd, ad, Wd, bd = x.double(), adj.double(), W.double(), bias.double()
        z0 = xd @ Wd[:dff].T + bd[:dff]
        z1 = xd @ Wd[dff:].T + bd[dff:]
        want = torch.relu(torch.diagonal(ad[:, 0], dim1=1, dim2=2).unsqueeze(-1) * z0 + ad[:, 1] @ z1).float()
        err = (y.cpu().reshape(B, K, dff) - want).abs().max().item() / want.abs().max().item()
        print(f"B={B} K={K} d={d} dff={dff} split={split}: rel err {err:.2e}", flush=True)
        assert err < 5e-5
print("ok")
