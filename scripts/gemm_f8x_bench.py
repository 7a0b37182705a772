"""ec_gemm_f16f8 (fp16 hi.hi + e4m3 cross terms) against ec_gemm_f16x3 (three fp16 products) on the ViT-B GEMM shapes of
the bench step (M = 32 images x 325 tokens), per tile mode: accuracy vs fp64 and algorithmic TFLOP/s."""
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
    return s.elapsed_time(e) / iters


def main():
    D = torch.device("cuda")
    M = 32 * 325
    for name, N, K in (("qkv", 2304, 768), ("proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)):
        x = torch.randn(M, K, device=D)
        w = torch.randn(N, K, device=D) * 0.02
        want = x.double() @ w.double().T
        err = lambda y: ((y.double() - want).abs().max() / want.abs().max()).item()
        fl = 2.0 * M * N * K
        y = torch.empty(M, N, device=D)
        line = f"{name:5s} M={M} N={N} K={K}:"
        for fmt, tag in ((ops.F16X2, "3xf16"), (ops.F16F8, "f16+2xf8")):
            a, b = ops.split_f16(x, fmt=fmt, role=0), ops.split_weight(w, fmt)
            for mode in (128, 256, 512):
                _lib.call("ec_tc_set_tile_n", mode)
                ops.gemm_tc(a, b, out=y)
                e = err(y)
                t = timeit(lambda: ops.gemm_tc(a, b, out=y))
                line += f" | {tag} m{mode} {t * 1e3:6.1f} us {fl / t / 1e9:4.0f} TF/s err {e:.0e}"
        _lib.call("ec_tc_set_tile_n", 0)
        print(line, flush=True)


if __name__ == "__main__":
    main()
