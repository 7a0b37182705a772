"""Offline (CPU) accuracy probe for a cheaper fp32-grade GEMM: keep a_hi*b_hi on fp16 MMAs, run the two small
cross terms a_lo*b_hi + a_hi*b_lo on fp8 (e4m3) MMAs at twice the rate -> 2 instead of 3 units of tensor time per
product.  Prints the error of each scheme against the fp64 product on a ViT-B shaped GEMM."""
import numpy as np
import torch

torch.manual_seed(0)
M, K, N = 512, 768, 768
x = torch.randn(M, K, dtype=torch.float64)
w = torch.randn(N, K, dtype=torch.float64) * 0.02
ref = x @ w.T


def split16(a):
    hi = a.to(torch.float16)
    return hi.double(), (a - hi.double()).to(torch.float16).double()


def q8(a):
    s = 2.0 ** np.floor(np.log2(240.0 / a.abs().max().item()))      # per-tensor power-of-two scale
    return (a * s).to(torch.float8_e4m3fn).double() / s


xh, xl = split16(x)
wh, wl = split16(w)
schemes = {
    "3 fp16 products (this round)": xl @ wh.T + xh @ wl.T + xh @ wh.T,
    "hi*hi only (plain fp16 GEMM)": xh @ wh.T,
    "fp16 hi*hi + fp8 cross terms": q8(xl) @ q8(wh).T + q8(xh) @ q8(wl).T + xh @ wh.T,
}
for name, y in schemes.items():
    print(f"{name:32s} max-rel {((y - ref).abs().max() / ref.abs().max()).item():.2e}  "
          f"rms-rel {((y - ref).norm() / ref.norm()).item():.2e}")
