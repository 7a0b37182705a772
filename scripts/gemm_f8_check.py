"""One pass over every F16F8 producer / consumer at shapes with partial tiles, against fp64 -- small enough to run under
compute-sanitizer (scripts/gpu_r03_z2.sh):  ec_split_f16f8 (roles 0 / 1), ec_gemm_f16f8 with the split-only TMA-store epilogue
(F16F8 and F16X2 rows out), the general epilogue (bias, LayerScale, residual, fp32 + split out), the CTA-pair tile mode, the
LayerNorm and the attention kernel writing F16F8 rows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from edgecape_b200 import ops

D = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def rnd(*s, scale=1.0):
    return (torch.randn(*s, generator=g) * scale).to(D)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def decode(sp, fmt, n):
    """fp32 value of split rows (hi16 + lo16, or hi16 + lo8 / 2^11)."""
    raw = sp.data.view(torch.uint8).reshape(-1, 4 * sp.Kp)
    hi = raw[:, :2 * sp.Kp].contiguous().view(torch.float16).float()
    if fmt == ops.F16X2:
        lo = raw[:, 2 * sp.Kp:].contiguous().view(torch.float16).float()
    else:
        e = raw[:, 2 * sp.Kp:].reshape(raw.shape[0], sp.Kp // 64, 2, 64)
        lo = e[:, :, 1].reshape(raw.shape[0], sp.Kp).contiguous().view(torch.float8_e4m3fn).float() / 2048.0
    return (hi + lo)[:, :n]


worst = 0.0
for (M, K, N) in [(650, 768, 2304), (1300, 3072, 768), (333, 200, 72), (2600, 768, 3072)]:
    x, w, b = rnd(M, K), rnd(N, K, scale=0.02), rnd(N, scale=0.1)
    ls, r = rnd(N, scale=0.5), rnd(M, N)
    a = ops.split_f16(x, fmt=ops.F16F8, role=0)
    wp = ops.split_weight(w, ops.F16F8)
    want = x.double() @ w.double().T + b.double()
    # split-only outputs through TMA stores, both row formats
    for fmt in (ops.F16F8, ops.F16X2):
        _, so = ops.gemm_tc(a, wp, bias=b, split_out=True, fp32_out=False, split_fmt=fmt)
        e = rel(decode(so, fmt, N), want)
        worst = max(worst, e)
        assert e < 2e-4, ("split-only", M, K, N, fmt, e)
    # general epilogue: LayerScale + residual, fp32 and split rows out
    got, so = ops.gemm_tc(a, wp, bias=b, colscale=ls, residual=r, split_out=True, split_fmt=ops.F16F8)
    want2 = want * ls.double() + r.double()
    e1, e2 = rel(got, want2), rel(decode(so, ops.F16F8, N), want2)
    worst = max(worst, e1)
    assert e1 < 5e-5 and e2 < 2e-4, ("general", M, K, N, e1, e2)
    # GELU epilogue -> F16F8 rows -> next GEMM (the fc1 -> fc2 chain)
    _, h2 = ops.gemm_tc(a, wp, bias=b, act=ops.ACT_GELU, split_out=True, fp32_out=False, split_fmt=ops.F16F8)
    w2 = rnd(64, N, scale=0.02)
    got = ops.gemm_tc(h2, ops.split_weight(w2, ops.F16F8))
    want3 = torch.nn.functional.gelu(want) @ w2.double().T
    e = rel(got, want3)
    worst = max(worst, e)
    assert e < 1e-4, ("chain", M, K, N, e)

# LayerNorm writing F16F8 rows, consumed by the GEMM
x, gam, bet = rnd(650, 768), rnd(768, scale=0.3) + 1.0, rnd(768, scale=0.1)
y2 = ops.layernorm(x, gam, bet, 1e-6, split="only", split_fmt=ops.F16F8)
want = torch.nn.functional.layer_norm(x.double(), (768,), gam.double(), bet.double(), 1e-6)
e = rel(decode(y2, ops.F16F8, 768), want)
assert e < 2e-4, ("layernorm", e)

# attention writing F16F8 rows
B, N, H, C = 2, 325, 12, 768
qkv = rnd(B * N, 3 * C)
qkv2 = ops.split_f16(qkv)
a2 = ops.attention_packed_split(qkv2, B, N, H, out_fmt=ops.F16F8)
q, k, v = (qkv.double().view(B, N, 3, H, 64)[:, :, i].transpose(1, 2) for i in range(3))
want = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v
want = want.transpose(1, 2).reshape(B * N, C)
e = rel(decode(a2, ops.F16F8, C), want)
assert e < 2e-4, ("attention", e)
torch.cuda.synchronize()
print("gemm_f8_check ok: worst fp32-output error %.2e, overflow counters %s" % (worst, ops.overflow_count()))
sys.exit(0)
