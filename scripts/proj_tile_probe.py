"""The proj GEMM of a ViT-B block (M = 32 x 325, N = K = 768, LayerScale + residual in place) in every tile mode and both
operand formats: 0 = the dispatcher's choice, 128 / 256 = single-CTA tiles, 512 = CTA pair."""
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(_HERE), _HERE]
from edgecape_b200 import _lib, ops  # noqa: E402
from gemm_f8_probe import timeit  # noqa: E402

D = torch.device("cuda")
M, C = 10400, 768
lib = _lib.load()
for fmt, tag in ((ops.F16X2, "3xf16"), (ops.F16F8, "f16+2xf8")):
    a2 = ops.split_f16(torch.randn(M, C, device=D), fmt=fmt)
    w = ops.split_f16(torch.randn(C, C, device=D) * 0.02, 1024.0, fmt=fmt, role=1)
    b, g, t = torch.randn(C, device=D), torch.rand(C, device=D), torch.randn(M, C, device=D)
    for tile in (0, 128, 256, 512):
        lib.ec_tc_set_tile_n(tile)
        us = timeit(lambda: ops.gemm_tc(a2, w, out=t, bias=b, colscale=g, residual=t))
        print(f"proj [{tag}] tile_n {tile:3d}: {us:6.1f} us", flush=True)
lib.ec_tc_set_tile_n(0)
