"""Fit and check of gelu_fast (csrc/common.cuh): erfc(t) = 2^(-t q(t)), q a degree-7 polynomial.
CPU only (numpy / scipy): prints the coefficients, the erf error of the fit in exact arithmetic, and the GELU error of
the fp32 evaluation order used on the device against the fp64 definition, with torch's fp32 GELU beside it."""
import numpy as np
import torch
from scipy.special import erf, erfc, log_ndtr

T, DEG = 4.6, 7


def q(t):
    return -(np.log(2.0) + log_ndtr(-t * np.sqrt(2.0))) / np.log(2.0) / t      # -log2(erfc(t)) / t


def fit():
    ts = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * T / 2 + T / 2
    w = erfc(ts) * ts                       # d erf = erfc(t) ln2 t dq: weight the residual accordingly
    V = np.vander(ts, DEG + 1, increasing=True)
    return np.linalg.lstsq(V * w[:, None], q(ts) * w, rcond=None)[0]


def gelu_fast(x, c32):
    x = x.astype(np.float32)
    t = np.minimum(np.abs(x) * np.float32(0.70710678118654752), np.float32(T)).astype(np.float32)
    qv = np.full_like(t, c32[DEG])
    for k in range(DEG - 1, -1, -1):
        qv = (qv * t + c32[k]).astype(np.float32)
    e = np.exp2(-(t * qv).astype(np.float32).astype(np.float64)).astype(np.float32)
    h = (np.float32(0.5) * x * e).astype(np.float32)
    return np.where(x >= 0, (x - h).astype(np.float32), h)


if __name__ == "__main__":
    coef = fit()
    tt = np.linspace(1e-6, T, 200001)
    print("erf error of the fit (exact arithmetic):", np.abs((1 - 2.0 ** (-tt * np.polyval(coef[::-1], tt))) - erf(tt)).max())
    c32 = coef.astype(np.float32)
    print("coefficients c0..c7:", [float(c) for c in c32])
    x = np.linspace(-12, 12, 2000001)
    want = 0.5 * x * (1 + erf(x / np.sqrt(2)))
    print("gelu_fast  max abs error vs fp64:", np.abs(gelu_fast(x, c32).astype(np.float64) - want).max())
    ref32 = torch.nn.functional.gelu(torch.from_numpy(x).float()).double().numpy()
    print("torch fp32 max abs error vs fp64:", np.abs(ref32 - want).max())
