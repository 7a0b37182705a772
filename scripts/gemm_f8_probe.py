"""Where does the F16F8 CTA-pair GEMM lose time?  The three large ViT-B GEMMs with their real epilogues (M = 32 x 325),
normal and with the kernel's experiment flags (results are wrong while a flag is set): 4 = no epilogue stores,
8 = no epilogue, 1 = operands stay resident in shared memory (no TMA after the first fill), 9 = neither."""
import sys

import torch

sys.path.insert(0, ".")
from edgecape_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def main():
    D = torch.device("cuda")
    M, C = 10400, 768
    lib = _lib.load()
    if len(sys.argv) > 2:
        lib.ec_tc_set_split_tma(int(sys.argv[2]))      # 1 = TMA stores (default), 2 = direct stores, 0 = transposing epilogue
    for fmt, tag in ((ops.F16F8, "f16+2xf8"), (ops.F16X2, "3xf16")):
        x2 = ops.split_f16(torch.randn(M, C, device=D), fmt=fmt)
        h2 = ops.split_f16(torch.randn(M, 4 * C, device=D), fmt=fmt)
        t = torch.randn(M, C, device=D)
        g = torch.rand(C, device=D)
        ws = {n: ops.split_f16(torch.randn(o, i, device=D) * 0.02, 1024.0, fmt=fmt, role=1) for n, (o, i) in
              dict(qkv=(3 * C, C), fc1=(4 * C, C), fc2=(C, 4 * C)).items()}
        bs = {n: torch.randn(w.rows, device=D) for n, w in ws.items()}
        cases = dict(
            qkv=lambda: ops.gemm_tc(x2, ws["qkv"], bias=bs["qkv"], split_out=True, fp32_out=False),
            fc1=lambda: ops.gemm_tc(x2, ws["fc1"], bias=bs["fc1"], act=ops.ACT_GELU, split_out=True, fp32_out=False,
                                    split_fmt=fmt),
            fc2=lambda: ops.gemm_tc(h2, ws["fc2"], out=t, bias=bs["fc2"], colscale=g, residual=t))
        lib.ec_tc_set_tile_n(512)
        for n, fn in cases.items():
            K = 4 * C if n == "fc2" else C
            line = f"[{tag} pair] {n}:"
            flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # one flag per process: a wrong experiment can hang
            lib.ec_tc_set_debug(flags)
            us = timeit(fn)
            line += f"  dbg{flags} {us:6.1f} us ({2.0 * M * ws[n].rows * K / us / 1e6:4.0f} TF/s)"
            lib.ec_tc_set_debug(0)
            print(line, flush=True)
            if flags == 0:      # where does the MMA thread wait?  (one traced launch)
                buf = torch.zeros(148, 4, dtype=torch.int64, device=D)
                lib.ec_tc_set_trace(buf.data_ptr())
                fn()
                torch.cuda.synchronize()
                lib.ec_tc_set_trace(None)
                tr = buf.cpu().double()
                tr = tr[tr[:, 3] > 0]
                print(f"      MMA thread of {len(tr)} leader CTAs: total {tr[:, 0].mean():.0f} clk, waiting for an accumulator "
                      f"(epilogue) {tr[:, 1].mean():.0f} ({100 * tr[:, 1].sum() / tr[:, 0].sum():.0f} %), waiting for operands (TMA) "
                      f"{tr[:, 2].mean():.0f} ({100 * tr[:, 2].sum() / tr[:, 0].sum():.0f} %), tiles {tr[:, 3].mean():.1f}", flush=True)
    lib.ec_tc_set_tile_n(0)


if __name__ == "__main__":
    main()
