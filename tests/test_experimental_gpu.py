"""EXPERIMENTAL kernels: compiled with the library but not on the default path.

The GPU test is skipped unless EDGECAPE_TEST_EXPERIMENTAL=1 -- the kernel was written after the round's GPU budget was
spent and has not run on hardware yet; the CPU test pins the operand format (plane scales) through the emulated ABI."""
import os

import numpy as np
import pytest
import torch

from edgecape_b200 import ops

from . import cpu_emulator


def _case(M=300, K=200, N=72, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    x[3, 5] = 37.0                                     # an outlier activation
    w = torch.randn(N, K, generator=g) * 0.02
    b = torch.randn(N, generator=g) * 0.1
    return x, w, b


def test_f16f8_operand_format_on_the_emulated_abi(monkeypatch):
    """fp16 hi.hi + two e4m3 cross terms with the static plane scales: ~1e-5 of the fp64 product, 30x better than
    the plain fp16 product."""
    cpu_emulator.install(monkeypatch)
    x, w, b = _case()
    got = ops.gemm_f16f8(x, w, bias=b, act=ops.ACT_RELU)
    want = torch.relu(x.double() @ w.double().T + b.double())
    err = ((got.double() - want).abs().max() / want.abs().max()).item()
    plain = torch.relu(x.half().double() @ w.half().double().T + b.double())
    err_plain = ((plain - want).abs().max() / want.abs().max()).item()
    assert err < 5e-5, err
    assert err < err_plain / 5, (err, err_plain)


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("EDGECAPE_TEST_EXPERIMENTAL") != "1",
                    reason="experimental kernel, not yet validated on hardware (set EDGECAPE_TEST_EXPERIMENTAL=1)")
@pytest.mark.parametrize("M,N,K", [(10400, 2304, 768), (300, 72, 200), (128, 256, 64), (1600, 768, 3072)])
def test_gemm_f16f8_matches_fp64(M, N, K):
    D = torch.device("cuda")
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.02
    b = torch.randn(N, generator=g) * 0.1
    got = ops.gemm_f16f8(x.to(D), w.to(D), bias=b.to(D)).cpu()
    want = x.double() @ w.double().T + b.double()
    err = ((got.double() - want).abs().max() / want.abs().max()).item()
    assert err < 1e-4, err
