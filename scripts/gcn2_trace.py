"""Phase trace of the project-first one-kernel GCN (ec_gcn_fused2_set_trace): median clock offsets of each phase of
a CTA's first item, over the CTAs."""
import sys

import torch

sys.path.insert(0, ".")
from edgecape_b200 import _lib, ops  # noqa: E402

NAMES = {0: "w:start", 1: "w:X block 0 stored", 2: "w:X block 1 stored", 3: "w:X block 2 stored", 4: "w:X block 3 stored",
         5: "w:A1 in TMEM (all warps)", 6: "w:T1 complete seen", 7: "w:T1 tiles written", 8: "w:D2 complete seen (epilogue starts)",
         9: "w:epilogue done", 11: "cta end", 12: "c:X block 0 seen", 13: "c:X block 1 seen", 14: "c:X block 2 seen",
         15: "c:X block 3 seen", 16: "c:GEMM A issued", 17: "c:T1 tiles seen", 18: "c:GEMM B issued", 10: "w:output tile read by the bulk stores",
         19: "w:raw block 0 arrived", 20: "w:block 0 in registers (barrier passed)", 21: "w:block 0 tiles written", 22: "w:block 0 proxy fence done"}


def main(B=64, K=100, d=256, dff=384, flags=0):
    D = torch.device("cuda")
    ops.TENSOR_CORES, ops.GCN_FUSED = True, 2
    x = torch.randn(B, K, d, device=D)
    adj = ops.soft_normalize_adj(torch.rand(B, K, K, device=D), torch.zeros(B, K, dtype=torch.uint8, device=D))
    Wp = ops.gcn_pack_weights(torch.randn(2 * dff, d, device=D) * d ** -0.5, torch.randn(2 * dff, device=D) * 0.1)
    out = torch.empty(B, K, dff, device=D)
    for _ in range(3):
        ops.gcn(x, adj, Wp, out=out)
    n = min(B * (dff // _lib.load().ec_gcn_fused2_slice(K, d, dff)), torch.cuda.get_device_properties(0).multi_processor_count)
    tr = torch.zeros(n, 32, dtype=torch.int64, device=D)
    _lib.call("ec_gcn_fused2_set_debug", flags)
    _lib.call("ec_gcn_fused2_set_trace", tr.data_ptr(), n)
    ops.gcn(x, adj, Wp, out=out)
    torch.cuda.synchronize()
    _lib.call("ec_gcn_fused2_set_trace", None, 0)
    _lib.call("ec_gcn_fused2_set_debug", 0)
    t = tr.cpu()
    rel = (t - t[:, :1]).double()
    print(f"B={B} K={K} d={d} dff={dff} dbg={flags}: {n} CTAs; median clocks after the workers' start")
    for i in sorted(NAMES, key=lambda i: rel[:, i].median().item()):
        if (t[:, i] == 0).all():
            continue
        print(f"  {rel[:, i].median().item():9.0f}  (min {rel[:, i].min().item():7.0f} max {rel[:, i].max().item():7.0f})  {NAMES[i]}")


if __name__ == "__main__":
    main()
    if "--experiments" in sys.argv:
        for flags in (1, 2, 4, 8, 12, 13):
            main(flags=flags)
    if "--short" not in sys.argv:
        main(B=16)
        main(B=2048)
