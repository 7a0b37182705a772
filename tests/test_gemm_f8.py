"""ec_gemm_f16f8 -- fp32-grade GEMM with a_hi.b_hi on fp16 and the two cross terms on e4m3 tensor cores -- and the
EC_SPLIT_F16F8 operand format its producers write (include/edgecape_b200.h).

CPU tests pin the format (plane scales, byte layout) and the host orchestration through the emulated ABI; the GPU
tests compare the kernel with fp64, with ec_gemm_f16x3 on the same inputs (every epilogue option, every tile mode),
and check the range behaviour (values beyond the e4m3 / fp16 range are counted, not silently wrong)."""
import os

import numpy as np
import pytest
import torch

import edgecape_b200 as E
from edgecape_b200 import ops
from edgecape_b200.config import state_dict_shapes
from edgecape_b200.synthetic import make_state_dict

from . import cpu_emulator


def _case(M=300, K=200, N=72, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    x[3, 5] = 37.0                                     # an outlier activation
    w = torch.randn(N, K, generator=g) * 0.02
    b = torch.randn(N, generator=g) * 0.1
    return x, w, b


def _f8_linear(x, w, **kw):
    a = ops.split_f16(x, fmt=ops.F16F8, role=0)
    return ops.gemm_tc(a, ops.split_weight(w, ops.F16F8), **kw)


def test_f16f8_operand_format_on_the_emulated_abi(monkeypatch):
    """fp16 hi.hi + two e4m3 cross terms with the static plane scales: ~1e-5 of the fp64 product, several times
    better than the plain fp16 product; the A-role planes written by the GEMM / LayerNorm epilogues equal the ones
    ec_split_f16f8 writes from the fp32 result."""
    cpu_emulator.install(monkeypatch)
    x, w, b = _case()
    got, so = _f8_linear(x, w, bias=b, act=ops.ACT_RELU, split_out=True, split_fmt=ops.F16F8)
    want = torch.relu(x.double() @ w.double().T + b.double())
    err = ((got.double() - want).abs().max() / want.abs().max()).item()
    plain = torch.relu(x.half().double() @ w.half().double().T + b.double())
    err_plain = ((plain - want).abs().max() / want.abs().max()).item()
    assert err < 5e-5, err
    assert err < err_plain / 5, (err, err_plain)
    assert so.fmt == ops.F16F8 and so.Kp == 128
    again = ops.split_f16(got, fmt=ops.F16F8, role=0)
    assert torch.equal(so.data.view(torch.uint8), again.data.view(torch.uint8))
    ln = ops.layernorm(x, torch.ones(200), torch.zeros(200), 1e-5, split="also", split_fmt=ops.F16F8)
    again = ops.split_f16(ln[0], fmt=ops.F16F8, role=0)
    assert torch.equal(ln[1].data.view(torch.uint8), again.data.view(torch.uint8))


def test_f16f8_byte_offsets_on_the_emulated_abi(monkeypatch):
    """The byte layout include/edgecape_b200.h documents: hi16 of column c at byte 2 c, hi8 at 2 Kp + 128 (c // 64) + c % 64,
    lo8 64 bytes behind it -- for both roles, checked entry by entry."""
    cpu_emulator.install(monkeypatch)
    x, w, _ = _case(M=5, K=200, N=3)
    for t, role, s_hi, s_lo in ((x, 0, 1.0, 2048.0), (w * 1024.0, 1, 2.0 ** -11, 1.0)):
        so = ops.split_f16(t, fmt=ops.F16F8, role=role)
        Kp = so.Kp
        assert Kp == 256
        raw = so.data.view(torch.uint8).reshape(t.shape[0], 4 * Kp)
        hi = t.half()
        lo = t - hi.float()
        for r in range(t.shape[0]):
            for c in (0, 1, 63, 64, 65, 127, 128, 199):
                off = 2 * Kp + 128 * (c // 64) + c % 64
                assert raw[r, 2 * c:2 * c + 2].view(torch.float16).item() == hi[r, c].item()
                assert raw[r, off:off + 1].view(torch.float8_e4m3fn).float().item() == \
                    (hi[r, c].float() * s_hi).to(torch.float8_e4m3fn).float().item()
                assert raw[r, off + 64:off + 65].view(torch.float8_e4m3fn).float().item() == \
                    (lo[r, c] * s_lo).to(torch.float8_e4m3fn).float().item()
            assert not raw[r, 2 * 200:2 * Kp].any()                      # zero padding of the hi16 plane
            assert not raw[r, 2 * Kp + 128 * 3 + 8:2 * Kp + 128 * 3 + 64].any()   # columns 200 .. 255 of the hi8 plane


def test_vit_host_orchestration_with_f8_linears(monkeypatch, golden_dir):
    """The ViT's F16F8 chain (LayerNorm -> qkv, LayerNorm -> fc1 -> fc2 on ec_gemm_f16f8; q, k, v / attention output /
    proj on F16X2) through the emulated ABI against the golden features of the unmodified reference."""
    from oracle.gen_golden import build_case
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(ops, "TENSOR_CORES", True)
    monkeypatch.setattr(ops, "f8_linear_ok", lambda M, w: True)
    calls = []
    orig = ops._lib.call
    monkeypatch.setattr(ops._lib, "call", lambda name, *a: (calls.append(name), orig(name, *a))[1])
    name = "tiny_k100_2shot_masked"
    golden = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    cfg, data, wseed = build_case(name)
    model = E.build_model(dict(model=cfg))
    model.load_state_dict(make_state_dict(state_dict_shapes(cfg), wseed), strict=True)
    model.eval()
    feat_q, _ = model.extract_features(data["img_s"], data["img_q"])
    depth = model.encoder_query.depth
    assert calls.count("ec_gemm_f16f8") == 3 * depth           # qkv, fc1, fc2 of every block
    a, w = feat_q.numpy()[:1], golden["feature_q"]
    err = np.abs(a.astype(np.float64) - w).max() / np.abs(w).max()
    assert err < 2e-4, err


# ------------------------------------------------------------------------------------------------ GPU
def _dev():
    return torch.device("cuda", torch.cuda.current_device())


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(10400, 2304, 768), (300, 72, 200), (128, 256, 64), (1600, 768, 3072), (1300, 768, 768),
                                   (517, 1000, 136)])
@pytest.mark.parametrize("tile_n", [0, 128, 256, 512])
def test_gemm_f16f8_matches_fp64(M, N, K, tile_n):
    D = _dev()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.02
    b = torch.randn(N, generator=g) * 0.1
    ops._lib.call("ec_tc_set_tile_n", tile_n)
    try:
        got = _f8_linear(x.to(D), w.to(D), bias=b.to(D)).cpu()
    finally:
        ops._lib.call("ec_tc_set_tile_n", 0)
    want = x.double() @ w.double().T + b.double()
    err = ((got.double() - want).abs().max() / want.abs().max()).item()
    assert err < 6e-5, err


@pytest.mark.gpu
@pytest.mark.parametrize("tile_n", [128, 512])
def test_gemm_f16f8_epilogues_match_f16x3(tile_n):
    """bias + GELU + LayerScale + residual into a strided 3-D view, and both split_out formats: the F16F8 kernel
    against ec_gemm_f16x3 on the same fp32 inputs."""
    D = _dev()
    g = torch.Generator().manual_seed(5)
    Bt, S, K, N = 9, 257, 384, 640
    x = torch.randn(Bt * S, K, generator=g).to(D)
    w = (torch.randn(N, K, generator=g) * 0.03).to(D)
    b = (torch.randn(N, generator=g) * 0.1).to(D)
    cs = torch.rand(N, generator=g).to(D)
    res = torch.randn(Bt * S, N, generator=g).to(D)
    buf3 = torch.zeros(Bt, S + 1, N, device=D)
    buf8 = torch.zeros(Bt, S + 1, N, device=D)
    ops._lib.call("ec_tc_set_tile_n", tile_n)
    try:
        ref, so3 = ops.gemm_tc(ops.split_f16(x), ops.split_weight(w), out=buf3[:, 1:, :], bias=b, act=ops.ACT_GELU,
                               colscale=cs, residual=res, split_out=True)
        for fmt in (ops.F16X2, ops.F16F8):
            got, so8 = _f8_linear(x, w, out=buf8[:, 1:, :], bias=b, act=ops.ACT_GELU, colscale=cs, residual=res,
                                  split_out=True, split_fmt=fmt)
            scale = ref.abs().max().item()
            assert (got - ref).abs().max().item() < 1e-4 * scale
            assert torch.equal(buf8[:, 0, :], torch.zeros_like(buf8[:, 0, :]))       # the skipped row is untouched
            flat = got.reshape(Bt * S, N).contiguous()
            want = ops.split_f16(flat, fmt=fmt, role=0)
            assert so8.fmt == fmt and torch.equal(so8.data.view(torch.uint8), want.data.view(torch.uint8))
    finally:
        ops._lib.call("ec_tc_set_tile_n", 0)


@pytest.mark.gpu
def test_f16f8_planes_reconstruct_and_layernorm_writes_them():
    D = _dev()
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(650, 768, generator=g) * 3).to(D)
    so = ops.split_f16(x, fmt=ops.F16F8, role=0)
    raw = so.data.view(torch.uint8).reshape(650, 4 * 768)
    hi16 = raw[:, :2 * 768].contiguous().view(torch.float16).float()
    e = raw[:, 2 * 768:].reshape(650, 768 // 64, 2, 64)          # per 64 columns: [hi8 x 64 | lo8 x 64]
    hi8 = e[:, :, 0].reshape(650, 768).contiguous().view(torch.float8_e4m3fn).float()
    lo8 = e[:, :, 1].reshape(650, 768).contiguous().view(torch.float8_e4m3fn).float()
    assert torch.equal(hi16, x.half().float())
    assert torch.equal(hi8, hi16.to(torch.float8_e4m3fn).float())
    assert ((hi16 + lo8 / 2048.0) - x).abs().max().item() <= 2.0 ** -15 * x.abs().max().item()
    w, b = torch.rand(768, generator=g).to(D) + 0.5, torch.randn(768, generator=g).to(D)
    y, so_ln = ops.layernorm(x, w, b, 1e-6, split="also", split_fmt=ops.F16F8)
    want = ops.split_f16(y, fmt=ops.F16F8, role=0)
    assert torch.equal(so_ln.data.view(torch.uint8), want.data.view(torch.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("C", [200, 72, 320])
def test_layernorm_scalar_path_writes_f16f8_rows(C):
    """Widths that are not a multiple of 128 take the scalar LayerNorm kernel: its F16F8 rows (values and the zero padding up
    to Kp) are the bytes ec_split_f16f8 writes from the fp32 result."""
    D = _dev()
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(77, C, generator=g) * 2).to(D)
    w, b = torch.rand(C, generator=g).to(D) + 0.5, torch.randn(C, generator=g).to(D)
    y, so = ops.layernorm(x, w, b, 1e-6, split="also", split_fmt=ops.F16F8)
    want = ops.split_f16(y, fmt=ops.F16F8, role=0)
    assert so.Kp == want.Kp == (C + 63) // 64 * 64
    assert torch.equal(so.data.view(torch.uint8), want.data.view(torch.uint8))
    ref = torch.nn.functional.layer_norm(x.double(), (C,), w.double(), b.double(), 1e-6)
    assert (y.double() - ref).abs().max().item() < 1e-5


def _interleave32(std, Kp):
    """role-1 rows [hi16 | per 64 columns: hi8, lo8] -> role-2 rows: 128 bytes per 32-column slice, [hi16 x 32 | hi8 x 32 | lo8 x 32]."""
    M = std.shape[0]
    h16 = std[:, :2 * Kp].reshape(M, Kp // 32, 64)
    e = std[:, 2 * Kp:].reshape(M, Kp // 64, 2, 64)              # role 1: per 64 columns [hi8 x 64 | lo8 x 64]
    h8 = e[:, :, 0].reshape(M, Kp // 32, 32)
    l8 = e[:, :, 1].reshape(M, Kp // 32, 32)
    return torch.cat((h16, h8, l8), dim=2).reshape(M, 4 * Kp)


def test_interleaved_weight_rows_on_the_emulated_abi(monkeypatch):
    """ec_split_f16f8 role 2 (the weight format of ec_gcn_fused2: planes interleaved per 32 columns) holds the bytes of
    role 1, re-ordered."""
    cpu_emulator.install(monkeypatch)
    _, w, _ = _case(K=200, N=72)
    std = ops.split_f16(w, 1024.0, fmt=ops.F16F8, role=1)
    il = ops.split_f16(w, 1024.0, fmt=ops.F16F8_I32, role=1)
    assert il.Kp == std.Kp
    a = std.data.view(torch.uint8).view(72, 4 * std.Kp)
    b = il.data.view(torch.uint8).view(72, 4 * std.Kp)
    assert torch.equal(b, _interleave32(a, std.Kp))


@pytest.mark.gpu
def test_interleaved_weight_rows_hold_the_same_bytes():
    """ec_split_f16f8 role 2 on the device: bit-identical to the role-1 rows re-ordered per 32-column slice."""
    w = (torch.randn(384, 516, generator=torch.Generator().manual_seed(5)) * 0.05).cuda()
    std = ops.split_f16(w, 4096.0, fmt=ops.F16F8, role=1)
    il = ops.split_f16(w, 4096.0, fmt=ops.F16F8_I32, role=1)
    a = std.data.view(torch.uint8).view(384, 4 * std.Kp)
    b = il.data.view(torch.uint8).view(384, 4 * std.Kp)
    assert torch.equal(b, _interleave32(a, std.Kp))


@pytest.mark.gpu
def test_f16f8_range_events_are_counted_and_degrade_gracefully():
    """An activation beyond 448 (e4m3 saturates) only loses its cross terms: the result stays within plain-fp16
    accuracy of that element's contribution and the event is counted; beyond 65504 hi16 itself overflows and the
    second counter says so.  A healthy run counts nothing."""
    D = _dev()
    g = torch.Generator().manual_seed(3)
    M, K, N = 512, 768, 512
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.02
    want = x.double() @ w.double().T
    ops.overflow_count(reset=True)
    got = _f8_linear(x.to(D), w.to(D)).cpu()
    assert ops.overflow_count() == (0, 0)
    assert ((got.double() - want).abs().max() / want.abs().max()).item() < 6e-5
    x2 = x.clone()
    x2[7, 11], x2[100, 5] = 500.0, -3000.0             # outlier channels of a real checkpoint's residual stream
    want2 = x2.double() @ w.double().T
    got2 = _f8_linear(x2.to(D), w.to(D)).cpu()
    n448, n65504 = ops.overflow_count()
    assert n448 >= 2 and n65504 == 0
    # the two rows with an outlier: error bounded by the outlier's lost cross terms (~2^-11 of its contribution)
    err_rows = (got2.double() - want2).abs().max(dim=1).values
    bound = 3000.0 * w.abs().max().item() * 2.0 ** -10
    assert err_rows.max().item() < bound
    clean = torch.ones(M, dtype=torch.bool)
    clean[[7, 100]] = False
    assert (err_rows[clean].max() / want2.abs().max()).item() < 6e-5
    x3 = x.clone()
    x3[1, 1] = 7e4
    _f8_linear(x3.to(D), w.to(D))
    assert ops.overflow_count(reset=True)[1] >= 1
    assert ops.overflow_count() == (0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("tile_n", [128, 256, 512])
@pytest.mark.parametrize("fmt_in", [0, 1], ids=["f16x3", "f16f8"])
@pytest.mark.parametrize("M,N,K", [(10400, 2304, 768), (1300, 1000, 320), (517, 136, 200), (128, 64, 64)])
def test_split_only_epilogue_through_tma_stores(M, N, K, fmt_in, tile_n):
    """GEMMs whose only output is split_out (qkv, fc1) assemble 128 x 64 blocks of the split rows in shared memory and
    TMA-store them.  The rows must equal -- bit for bit -- the split form of the fp32 result of the same GEMM, in both
    output formats, including the M tail (rows clipped by the tensor map) and the zero padding of columns >= N."""
    D = _dev()
    g = torch.Generator().manual_seed(M + N + K + tile_n)
    x = torch.randn(M, K, generator=g).to(D)
    w = (torch.randn(N, K, generator=g) * 0.05).to(D)
    b = (torch.randn(N, generator=g) * 0.1).to(D)
    a2 = ops.split_f16(x, fmt=fmt_in, role=0)
    b2 = ops.split_weight(w, fmt_in)
    ops._lib.call("ec_tc_set_tile_n", tile_n)
    try:
        ref = ops.gemm_tc(a2, b2, bias=b, act=ops.ACT_GELU)
        for fmt_out in ([ops.F16X2, ops.F16F8] if fmt_in == ops.F16F8 else [ops.F16X2]):
            so = ops.gemm_tc(a2, b2, bias=b, act=ops.ACT_GELU, split_out=True, fp32_out=False, split_fmt=fmt_out)[1]
            want = ops.split_f16(ref, fmt=fmt_out, role=0)
            assert so.Kp == want.Kp and torch.equal(so.data.view(torch.uint8), want.data.view(torch.uint8)), fmt_out
            for mode in (0, 2):                                     # the other two epilogue forms give the same bytes
                ops._lib.call("ec_tc_set_split_tma", mode)
                so0 = ops.gemm_tc(a2, b2, bias=b, act=ops.ACT_GELU, split_out=True, fp32_out=False, split_fmt=fmt_out)[1]
                ops._lib.call("ec_tc_set_split_tma", 1)
                assert torch.equal(so0.data.view(torch.uint8), want.data.view(torch.uint8)), mode
    finally:
        ops._lib.call("ec_tc_set_split_tma", 1)
        ops._lib.call("ec_tc_set_tile_n", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,fmt", [(10400, 768, 3072, 1), (10400, 768, 768, 0), (10400, 768, 768, 1), (5200, 768, 3072, 1)],
                         ids=["fc2_pair_f8", "proj_f16x3", "proj_f8", "fc2_batch8"])
def test_gemm_ksplit_tail(M, N, K, fmt):
    """ec_tc_set_ksplit(1): the tiles of a partly empty last wave are cut into three K-parts that accumulate onto one
    another IN A FIXED ORDER through C (residual in place, LayerScale, bias): close to fp64, to the unsplit launch, and
    bit-identical from run to run."""
    from edgecape_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * 0.02).cuda()
    b = (torch.randn(N, generator=g) * 0.1).cuda()
    ls = torch.rand(N, generator=g).cuda()
    t0 = torch.randn(M, N, generator=g).cuda()
    a = ops.split_f16(x, fmt=fmt, role=0)
    ws = ops.split_weight(w, fmt)
    want = (t0.double() + ls.double() * (x.double() @ w.double().T + b.double())).float()

    def run():
        t = t0.clone()
        ops.gemm_tc(a, ws, out=t, bias=b, colscale=ls, residual=t)
        return t
    ref = run()
    n0 = lib.ec_tc_ksplit_launches()
    lib.ec_tc_set_ksplit(1)
    try:
        y1, y2 = run(), run()
    finally:
        lib.ec_tc_set_ksplit(0)
    assert lib.ec_tc_ksplit_launches() - n0 == 2, "the K-split did not engage at this shape"
    assert torch.equal(y1, y2)
    scale = want.abs().max().item()
    assert (y1 - want).abs().max().item() / scale < 3e-5
    assert (y1 - ref).abs().max().item() / scale < 2e-5
