"""Phase trace of the one-kernel GCN (ec_gcn_fused_set_trace): median clock offsets of each phase over the CTAs."""
import sys

import torch

sys.path.insert(0, ".")
from edgecape_b200 import _lib, ops  # noqa: E402

NAMES = {0: "w:start", 1: "w:rowsum done", 2: "w:A1 tiles done", 3: "w:X tiles done (FILL)", 4: "w:GEMM1 done (D1)",
         5: "w:a0 rescale done (ZA)", 6: "w:ZB0", 7: "w:ZB1", 8: "w:ZB2", 9: "w:ZB3", 10: "w:ACC ready", 11: "w:end",
         12: "c:FILL seen", 13: "c:GEMM1 issued", 14: "c:ZA seen"}
for kb in range(8):
    NAMES[15 + 2 * kb] = f"c:kb{kb} operands ready"
    NAMES[16 + 2 * kb] = f"c:kb{kb} issued"


def main(B=64, K=100, d=256, dff=384):
    D = torch.device("cuda")
    ops.TENSOR_CORES, ops.GCN_FUSED = True, True
    x = torch.randn(B, K, d, device=D)
    adj = ops.soft_normalize_adj(torch.rand(B, K, K, device=D), torch.zeros(B, K, dtype=torch.uint8, device=D))
    Wp = ops.gcn_pack_weights(torch.randn(2 * dff, d, device=D) * d ** -0.5, torch.randn(2 * dff, device=D) * 0.1)
    out = torch.empty(B, K, dff, device=D)
    for _ in range(3):
        ops.gcn(x, adj, Wp, out=out)
    n = B * (dff // _lib.load().ec_gcn_fused_slice(K, d, dff))
    tr = torch.zeros(n, 32, dtype=torch.int64, device=D)
    _lib.call("ec_gcn_fused_set_trace", tr.data_ptr(), n)
    ops.gcn(x, adj, Wp, out=out)
    torch.cuda.synchronize()
    _lib.call("ec_gcn_fused_set_trace", None, 0)
    t = tr.cpu()
    rel = (t - t[:, :1]).double()
    print(f"B={B} K={K} d={d} dff={dff}: {n} CTAs; median clocks after the workers' start")
    order = sorted(NAMES, key=lambda i: rel[:, i].median().item())
    for i in order:
        if (t[:, i] == 0).all():
            continue
        print(f"  {rel[:, i].median().item():9.0f}  (min {rel[:, i].min().item():7.0f} max {rel[:, i].max().item():7.0f})  {NAMES[i]}")


if __name__ == "__main__":
    main()
    if "--experiments" in sys.argv:
        for flags, what in ((1, "no global loads of A1 / X"), (2, "no epilogue stores"), (3, "neither")):
            print(f"--- experiment: {what}")
            _lib.call("ec_gcn_fused_set_debug", flags)
            main()
        _lib.call("ec_gcn_fused_set_debug", 0)
    main(B=16)
