"""EXPERIMENTAL: ec_gemm_f16f8 (fp16 hi.hi + e4m3 cross terms) against ec_gemm_f16x3 (three fp16 products) on the ViT-B
GEMM shapes of the bench step (M = 32 images x 325 tokens): accuracy vs fp64 and algorithmic TFLOP/s.
First thing to run on a B200 next round:  EDGECAPE_TEST_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -m gpu
then  python scripts/gemm_f8x_bench.py"""
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
        a2, b2 = ops.split_f16(x), ops.split_weight(w)
        y3 = ops.gemm_tc(a2, b2)
        a3, Kp = ops.split_f16f8(x, 1.0, 0)
        scale = b2.scale
        b3, _ = ops.split_f16f8(w, scale, 1)
        y8 = torch.empty(M, N, device=D)
        run8 = lambda: _lib.call("ec_gemm_f16f8", a3.data_ptr(), b3.data_ptr(), y8.data_ptr(), M, N, Kp, N, 1.0 / scale, None, 0,
                                 torch.cuda.current_stream().cuda_stream)
        run8()
        err = lambda y: ((y.double() - want).abs().max() / want.abs().max()).item()
        t3 = timeit(lambda: ops.gemm_tc(a2, b2, out=y3))
        _lib.call("ec_tc_set_tile_n", 256)          # the same 128x256 single-CTA tile the f16f8 kernel uses
        t3s = timeit(lambda: ops.gemm_tc(a2, b2, out=y3))
        _lib.call("ec_tc_set_tile_n", 0)
        t8 = timeit(run8)
        fl = 2.0 * M * N * K
        print(f"{name:5s} M={M} N={N} K={K}: 3xfp16 auto {t3 * 1e3:7.1f} us {fl / t3 / 1e9:6.0f} TF/s err {err(y3):.1e} | "
              f"3xfp16 128x256 {t3s * 1e3:7.1f} us {fl / t3s / 1e9:6.0f} TF/s | "
              f"fp16+2xfp8 128x256 {t8 * 1e3:7.1f} us {fl / t8 / 1e9:6.0f} TF/s err {err(y8):.1e}", flush=True)


if __name__ == "__main__":
    main()
