"""Tensor-level wrappers over the C ABI (include/edgecape_b200.h).

Every function takes CUDA fp32 torch tensors (views with a unit innermost stride are fine: row
and batch strides are forwarded as ld / stride arguments), enqueues the kernels on torch's
current stream and returns the output tensor.  PyTorch is used for allocation and streams only;
no torch compute op is called here, and nothing falls back to the CPU.
"""
import math
import os
import weakref

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH = 0, 1, 2, 3
RES_NONE, RES_ADD, RES_GATE = 0, 1, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, name, dtype=torch.float32):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.EdgeCapeLibraryError(
            f"{name} is on {t.device}: edgecape_b200 has no CPU path (the CUDA library is the product)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's current stream (_stream); a tensor living elsewhere would be
        # dereferenced on the wrong GPU
        raise _lib.EdgeCapeLibraryError(
            f"{name} lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}: "
            f"call torch.cuda.set_device({t.device.index}) (one device per process is the supported layout)")


def _rows(t, name):
    """2-D / 3-D view with unit inner stride -> (batch, rows, cols, ld, batch_stride)."""
    _chk(t, name)
    if t.dim() == 2:
        M, C = t.shape
        assert C <= 1 or t.stride(1) == 1, f"{name}: inner stride must be 1"
        return 1, M, C, (t.stride(0) if M > 1 else max(C, t.stride(0))), 0
    assert t.dim() == 3, f"{name}: expected 2-D or 3-D"
    B, M, C = t.shape
    assert C <= 1 or t.stride(2) == 1, f"{name}: inner stride must be 1"
    ld = t.stride(1) if M > 1 else max(C, t.stride(1))
    return B, M, C, ld, (t.stride(0) if B > 1 else 0)


def empty(*shape, dtype=torch.float32, device=None):
    return torch.empty(*shape, dtype=dtype, device=device or torch.device("cuda", torch.cuda.current_device()))


def gemm(a, b, out=None, b_kmajor=True, bias=None, act=ACT_NONE, colscale=None, residual=None,
         res_mode=RES_ADD):
    """out = epilogue(a @ b^T) (b_kmajor: b is [N,K]) or a @ b (b is [K,N]); 2-D or batched 3-D."""
    ba, M, K, lda, sA = _rows(a, "a")
    bb, r0, r1, ldb, sB = _rows(b, "b")
    N, Kb = (r0, r1) if b_kmajor else (r1, r0)
    assert Kb == K, f"gemm: inner dimensions differ ({K} vs {Kb})"
    batch = max(ba, bb)
    assert ba in (1, batch) and bb in (1, batch)
    if out is None:
        out = empty(batch, M, N, device=a.device) if (a.dim() == 3 or b.dim() == 3) else empty(M, N, device=a.device)
    bo, Mo, No, ldc, sC = _rows(out, "out")
    assert (Mo, No) == (M, N) and bo == batch, f"gemm: out shape {tuple(out.shape)} != {(batch, M, N)}"
    ldr, sR = 0, 0
    if residual is not None:
        br, Mr, Nr, ldr, sR = _rows(residual, "residual")
        assert (Mr, Nr) == (M, N) and br in (1, batch)
    else:
        res_mode = RES_NONE
    _chk(bias, "bias")
    _chk(colscale, "colscale")
    _lib.call("ec_gemm", _p(a), _p(b), _p(out), M, N, K, lda, ldb, ldc, 1 if b_kmajor else 0, batch,
              sA if ba > 1 else 0, sB if bb > 1 else 0, sC, _p(bias), act, _p(colscale), _p(residual), ldr, sR,
              res_mode, _stream())
    return out


# ------------------------------------------------------------- tensor-core (tcgen05) linears
# EDGECAPE_TC=0 routes every linear through the fp32 SIMT GEMM (bit-faithful FFMA reference path);
# the default uses the split-fp16 tcgen05 GEMM wherever the shape fills a 128x128 tile reasonably.
TENSOR_CORES = os.environ.get("EDGECAPE_TC", "1") != "0"
ATTENTION_TC = os.environ.get("EDGECAPE_ATTN_TC", "1") != "0"     # tcgen05 attention for head dim 64
ATTENTION_TMA = os.environ.get("EDGECAPE_ATTN_TMA", "1") != "0"   # TMA-fed variant on pre-split QKV (ViT)
TC_MIN_M, TC_MIN_N, TC_MIN_K = 64, 32, 32
# Split-operand row formats (include/edgecape_b200.h): F16X2 = [hi16 | lo16] for ec_gemm_f16x3 and the attention
# kernels, F16F8 = [hi16 | hi8, lo8 interleaved per 64 columns] for ec_gemm_f16f8 (a_hi.b_hi on fp16, both cross terms on e4m3: 2 instead of 3
# units of tensor time).  EDGECAPE_GEMM_F8=0 keeps every linear on three fp16 products; with it on (default) the
# large linears (M >= F8_MIN_M rows: the ViT's qkv / fc1 / fc2 at bench batch sizes) take the F16F8 kernel.
F16X2, F16F8 = 0, 1
F16F8_I32 = 2      # weights only: the F16F8 B-role planes interleaved per 32 columns, [hi16 x 32 | hi8 x 32 | lo8 x 32] (the fused
                   # GCN streams its weights in 32-deep k-slices: one TMA box with 128-byte rows then holds a whole slice)
GEMM_F8 = os.environ.get("EDGECAPE_GEMM_F8", "1") != "0"
# EDGECAPE_PROJ_F8=1: the ViT attention writes F16F8 rows and the proj GEMM runs on ec_gemm_f16f8 too.  Off by default: proj
# (K = 768, 128x128 tiles, 3.3 waves) is bound by per-tile latency, not by tensor time -- measured 4.20 vs 4.23 ms of GEMM
# time per step, no change of the step (profiles/r03_h_bench_proj_f8_{0,1}.json) -- so it keeps three fp16 products.
PROJ_F8 = os.environ.get("EDGECAPE_PROJ_F8", "0") != "0"
F8_MIN_M = 2048
_SPLIT_WEIGHTS = {}


class SplitOperand:
    """Split form of an fp32 matrix (scaled by `scale`, a power of two): fp16 [rows, 2*Kp] holding 4*Kp bytes per
    row in the format `fmt` (F16X2: [hi | lo] halves; F16F8: [hi16 | per 64 columns: hi8, lo8])."""
    __slots__ = ("data", "rows", "K", "Kp", "scale", "fmt")

    def __init__(self, data, rows, K, Kp, scale, fmt=F16X2):
        self.data, self.rows, self.K, self.Kp, self.scale, self.fmt = data, rows, K, Kp, scale, fmt


def _kp(K):
    return (K + 63) // 64 * 64


def split_f16(x, scale=1.0, out=None, fmt=F16X2, role=0):
    """fp32 rows (2-D view, or 3-D [B,S,K] view read in place) -> SplitOperand in the row format `fmt`
    (role: 0 = A operand / activations, 1 = B operand / weights -- the F16F8 plane scales differ)."""
    M, K, ldx, seg, seg_stride = _seg(x, "x")
    Kp = _kp(K)
    if out is None:
        out = empty(M, 2 * Kp, dtype=torch.float16, device=x.device)
    assert out.is_contiguous() and out.dtype == torch.float16 and out.numel() == M * 2 * Kp
    if fmt == F16X2:
        _lib.call("ec_split_f16", _p(x), _p(out), M, K, ldx, seg, seg_stride, Kp, float(scale), _stream())
    elif fmt == F16F8_I32:
        assert role == 1, "the interleaved e4m3 planes are a weight (B operand) format"
        _lib.call("ec_split_f16f8", _p(x), _p(out), M, K, ldx, seg, seg_stride, Kp, float(scale), 2, _stream())
    else:
        _lib.call("ec_split_f16f8", _p(x), _p(out), M, K, ldx, seg, seg_stride, Kp, float(scale), int(role), _stream())
    return SplitOperand(out, M, K, Kp, float(scale), fmt)


def overflow_count(reset=False):
    """{values beyond the e4m3 range (448), values beyond the fp16 range (65504)} seen so far by the F16F8 split
    producers on the current device.  Synchronises the device (a diagnostic, not part of the hot path)."""
    import ctypes
    buf = (ctypes.c_ulonglong * 2)()
    _lib.call("ec_overflow_count", ctypes.cast(buf, ctypes.c_void_p), int(bool(reset)))
    return int(buf[0]), int(buf[1])


def clear_weight_cache():
    """Drop the cached split copies of weight matrices.  The cache is keyed on the parameter's identity, address and
    `_version`, which an in-place update through `.data` does NOT bump: the detector calls this from
    `load_state_dict` / `.to()` / `.cuda()`, and code that writes weights through `.data` must call it too."""
    _SPLIT_WEIGHTS.clear()


def split_weight(w, fmt=F16X2):
    """Cached split form of a weight matrix [N,K] (one-time repack per checkpoint: the power-of-two
    scale is picked from the tensor's absmax so that small weights stay out of the fp16 subnormals)."""
    owner = w._base if w._base is not None else w        # the nn.Parameter behind a reshaped view
    key = (id(owner), w.data_ptr(), owner._version, tuple(w.shape), fmt)
    hit = _SPLIT_WEIGHTS.get(key)
    sw = hit[1] if hit is not None and hit[0]() is owner else None   # guards against id / address reuse
    if sw is None:
        amax = float(w.detach().abs().max())
        scale = 1.0
        if amax > 0 and math.isfinite(amax):
            scale = 2.0 ** max(-8, min(14, math.floor(math.log2(16384.0 / amax))))
        sw = split_f16(w.detach(), scale, fmt=fmt, role=1)
        if len(_SPLIT_WEIGHTS) > 4096:
            _SPLIT_WEIGHTS.clear()
        _SPLIT_WEIGHTS[key] = (weakref.ref(owner), sw)
    return sw


def gemm_tc(a2, b2, out=None, bias=None, act=ACT_NONE, colscale=None, residual=None, res_mode=RES_ADD,
            split_out=False, fp32_out=True, res_rows=0, split_fmt=F16X2):
    """out = epilogue(A @ B^T) on the tcgen05 tensor cores from two SplitOperands of the same row format
    (F16X2 -> ec_gemm_f16x3, F16F8 -> ec_gemm_f16f8).  `out` may be a 2-D view or a 3-D [B,S,N] view with
    B*S == M (batch-strided rows).  With split_out=True also returns the SplitOperand of the result (produced by
    the epilogue, in the row format `split_fmt`)."""
    assert a2.Kp == b2.Kp, f"gemm_tc: padded K differs ({a2.Kp} vs {b2.Kp})"
    assert a2.fmt == b2.fmt, "gemm_tc: operands must share a split format"
    assert split_fmt == F16X2 or a2.fmt == F16F8, "only ec_gemm_f16f8 writes F16F8 rows"
    M, N = a2.rows, b2.rows
    dev = a2.data.device
    ldc = seg_c = seg_stride_c = 0
    if fp32_out:
        if out is None:
            out = empty(M, N, device=dev)
        Mo, No, ldc, seg_c, seg_stride_c = _seg(out, "out")
        assert (Mo, No) == (M, N), f"gemm_tc: out {tuple(out.shape)} does not hold {(M, N)}"
    else:
        assert split_out and out is None, "fp32_out=False only makes sense with split_out=True"
    ldr = 0
    if residual is not None:
        Mr, Nr, ldr, seg_r, _ = _seg(residual, "residual")
        assert (Mr, Nr) == (res_rows or M, N) and seg_r == 0, "gemm_tc: residual must be a plain 2-D row view"
    else:
        res_mode = RES_NONE
    _chk(bias, "bias")
    _chk(colscale, "colscale")
    so, so_ptr, so_kp = None, None, 0
    if split_out:
        so_kp = _kp(N)
        buf = (torch.zeros if so_kp != N else torch.empty)(M, 2 * so_kp, dtype=torch.float16, device=dev)
        so = SplitOperand(buf, M, N, so_kp, 1.0, split_fmt)
        so_ptr = buf.data_ptr()
    if a2.fmt == F16F8:
        _lib.call("ec_gemm_f16f8", _p(a2.data), _p(b2.data), _p(out), M, N, a2.Kp, ldc, seg_c, seg_stride_c,
                  1.0 / (a2.scale * b2.scale), _p(bias), act, _p(colscale), _p(residual), ldr, res_mode, res_rows,
                  so_ptr, so_kp, 1.0, split_fmt, _stream())
    else:
        _lib.call("ec_gemm_f16x3", _p(a2.data), _p(b2.data), _p(out), M, N, a2.Kp, ldc, seg_c, seg_stride_c,
                  1.0 / (a2.scale * b2.scale), _p(bias), act, _p(colscale), _p(residual), ldr, res_mode, res_rows,
                  so_ptr, so_kp, 1.0, _stream())
    return (out, so) if split_out else out


def f8_linear_ok(M, w):
    """Would a linear with `M` activation rows and weight `w` [N,K] take the F16F8 kernel?  Callers that produce its A
    operand (LayerNorm, the previous GEMM's epilogue) ask before choosing the row format."""
    if w.dim() > 2:
        w = w.reshape(w.shape[0], -1)
    return bool(TENSOR_CORES and GEMM_F8 and M >= F8_MIN_M and w.shape[0] >= 256 and w.shape[1] >= 256)


def _tc_ok(x, w, residual):
    if not TENSOR_CORES:
        return False
    if residual is not None and residual.dim() == 3 and _seg(residual, "residual")[3] != 0:
        return False
    if isinstance(x, SplitOperand):
        return True
    M = x.numel() // x.shape[-1]
    return M >= TC_MIN_M and w.shape[0] >= TC_MIN_N and w.shape[1] >= TC_MIN_K


def linear(x, w, bias=None, out=None, act=ACT_NONE, colscale=None, residual=None, res_mode=RES_ADD,
           split_out=False, fp32_out=True, split_fmt=F16X2):
    """nn.Linear on the last dimension of a 2-D/3-D view (or an already split A operand); w is [N,K]
    (conv 1x1 weights are reshaped).  Dispatches to a tcgen05 kernel (the one matching the split format of `x`) or
    the fp32 SIMT kernel."""
    if w.dim() > 2:
        w = w.reshape(w.shape[0], -1)
    if _tc_ok(x, w, residual):
        if isinstance(x, SplitOperand):
            a2 = x
        else:   # fp32 rows: split here, in the format of the kernel this shape takes
            a2 = split_f16(x, fmt=F16F8 if f8_linear_ok(x.numel() // x.shape[-1], w) else F16X2)
        if out is None and not isinstance(x, SplitOperand) and x.dim() == 3:
            out = empty(x.shape[0], x.shape[1], w.shape[0], device=x.device)
        if not fp32_out:
            out = None
        return gemm_tc(a2, split_weight(w, a2.fmt), out=out, bias=bias, act=act, colscale=colscale, residual=residual,
                       res_mode=res_mode, split_out=split_out, fp32_out=fp32_out, split_fmt=split_fmt)
    assert not isinstance(x, SplitOperand), "a split operand needs the tensor-core path"
    y = gemm(x, w, out=out, b_kmajor=True, bias=bias, act=act, colscale=colscale, residual=residual,
             res_mode=res_mode)
    return (y, None) if split_out else y


def tc_linear_ok(x, w):
    """Would linear(x, w) run on the tcgen05 kernel?  (callers use it to choose the split-fp16 chain)"""
    if w.dim() > 2:
        w = w.reshape(w.shape[0], -1)
    return _tc_ok(x, w, None)


def linear_split(x, w, bias=None, act=ACT_NONE):
    """linear() whose only output is the split-fp16 operand the next tensor-core kernel consumes."""
    assert tc_linear_ok(x, w)
    return linear(x, w, bias, act=act, split_out=True, fp32_out=False)[1]


def layernorm(x, w, b, eps=1e-5, out=None, residual=None, sum_out=None, split="no", split_fmt=F16X2):
    """LayerNorm over the last dim of x (+ residual).  A 3-D x view [B,S,C] with a batch stride
    larger than S*ld (e.g. ViT tokens without the cls row) is read in place.
    split: "no" -> fp32 result; "also" -> (fp32, SplitOperand); "only" -> SplitOperand (the fp32
    result is never written: the consumer is a tensor-core GEMM)."""
    _chk(x, "x")
    seg, seg_stride = 0, 0
    if x.dim() == 3:
        B, S, C = x.shape
        ldx = x.stride(1)
        if x.stride(0) != S * ldx:
            seg, seg_stride = S, x.stride(0)
        M = B * S
    else:
        M, C = x.shape
        ldx = x.stride(0)
    assert x.stride(-1) == 1
    o2, ldy = None, 0
    if split != "only":
        if out is None:
            out = empty(*x.shape, device=x.device)
        o2 = out.reshape(M, C) if out.is_contiguous() else out
        assert o2.dim() == 2 and o2.stride(1) == 1
        ldy = o2.stride(0)
    so, so_ptr, so_kp = None, None, 0
    if split != "no":
        so_kp = _kp(C)
        so = SplitOperand(empty(M, 2 * so_kp, dtype=torch.float16, device=x.device), M, C, so_kp, 1.0, split_fmt)
        so_ptr = so.data.data_ptr()
    ldr = ld_sum = 0
    if residual is not None:
        _chk(residual, "residual")
        r2 = residual.reshape(M, C) if residual.is_contiguous() else residual
        assert r2.dim() == 2 and r2.stride(1) == 1
        residual, ldr = r2, r2.stride(0)
    if sum_out is not None:
        s2 = sum_out.reshape(M, C) if sum_out.is_contiguous() else sum_out
        assert s2.dim() == 2 and s2.stride(1) == 1
        sum_out, ld_sum = s2, s2.stride(0)
    _lib.call("ec_layernorm", _p(x), ldx, seg, seg_stride, _p(residual), ldr, _p(sum_out), ld_sum, _p(o2),
              ldy, _p(w), _p(b), float(eps), M, C, so_ptr, so_kp, split_fmt, _stream())
    return so if split == "only" else ((out, so) if split == "also" else out)


def add_rows_(x, pos, S):
    """x[b, :S, :] += pos[:S, :]  in place; x contiguous [B,T,C]."""
    _chk(x, "x"); _chk(pos, "pos")
    assert x.is_contiguous() and pos.is_contiguous()
    B, T, C = x.shape
    _lib.call("ec_add_rows", _p(x), _p(pos), B, T, S, C, _stream())
    return x


def _seg(t, name):
    """2-D [M,C] or 3-D [B,S,C] view -> (M, C, ld, seg, seg_stride)."""
    _chk(t, name)
    assert t.shape[-1] <= 1 or t.stride(-1) == 1, f"{name}: inner stride must be 1"
    if t.dim() == 2:
        return t.shape[0], t.shape[1], t.stride(0), 0, 0
    B, S, C = t.shape
    if t.stride(0) == S * t.stride(1):
        return B * S, C, t.stride(1), 0, 0
    return B * S, C, t.stride(1), S, t.stride(0)


def copy_rows(x, out, bcast_rows=0):
    """out[m, :] = x[m % bcast_rows if bcast_rows else m, :] for 2-D / 3-D row views."""
    Mx, C, ldx, sx, ssx = _seg(x, "x")
    M, Co, ldy, sy, ssy = _seg(out, "out")
    assert C == Co and (bcast_rows > 0 or Mx == M)
    _lib.call("ec_copy_rows", _p(x), ldx, sx, ssx, _p(out), ldy, sy, ssy, M, C, bcast_rows, _stream())
    return out


def axpby(x, y, a=1.0, b=1.0, div=1.0, out=None):
    """(a*x + b*y) / div elementwise on contiguous tensors of equal size."""
    _chk(x, "x"); _chk(y, "y")
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
    if out is None:
        out = empty(*x.shape, device=x.device)
    _lib.call("ec_axpby", _p(x), _p(y), _p(out), float(a), float(b), float(div), x.numel(), _stream())
    return out


def attention(q, k, v, nheads, scale=None, key_mask=None, bias=None, out=None, split="no"):
    """q [B,Lq,H*D], k [B,Lk,H*D], v [B,Lk,H*D] views (unit inner stride) -> [B,Lq,H*D].
    split as in layernorm(): "only" returns just the SplitOperand [B*Lq, 2*H*D] of the output."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, n)
        assert t.dim() == 3 and t.stride(2) == 1
    B, Lq, E = q.shape
    Lk = k.shape[1]
    D = E // nheads
    assert k.shape[2] == E and v.shape[2] == E and v.shape[1] == Lk
    ldo = so_ = 0
    if split != "only":
        if out is None:
            out = empty(B, Lq, E, device=q.device)
        assert out.stride(2) == 1
        ldo, so_ = out.stride(1), out.stride(0)
    else:
        out = None
    sp, sp_ptr = None, None
    if split != "no":
        assert E % 64 == 0, "split attention output needs H*D to be a multiple of 64"
        sp = SplitOperand(empty(B * Lq, 2 * E, dtype=torch.float16, device=q.device), B * Lq, E, E, 1.0)
        sp_ptr = sp.data.data_ptr()
    if scale is None:
        scale = D ** -0.5
    _chk(key_mask, "key_mask", torch.uint8)
    _chk(bias, "bias")
    if key_mask is not None:
        assert key_mask.is_contiguous() and tuple(key_mask.shape) == (B, Lk)
    if bias is not None:
        assert bias.is_contiguous() and tuple(bias.shape) == (B, nheads, Lq, Lk)
    if TENSOR_CORES and ATTENTION_TC and D in (32, 64) and Lk <= 448:
        _lib.call("ec_attention_tc", _p(q), _p(k), _p(v), _p(out), B, nheads, Lq, Lk, D, q.stride(1), k.stride(1),
                  v.stride(1), ldo, q.stride(0), k.stride(0), v.stride(0), so_, float(scale), _p(key_mask), _p(bias),
                  sp_ptr, E if sp is not None else 0, _stream())
        return sp if split == "only" else ((out, sp) if split == "also" else out)
    _lib.call("ec_attention", _p(q), _p(k), _p(v), _p(out), B, nheads, Lq, Lk, D, q.stride(1), k.stride(1),
              v.stride(1), ldo, q.stride(0), k.stride(0), v.stride(0), so_, float(scale),
              _p(key_mask), _p(bias), sp_ptr, E if sp is not None else 0, _stream())
    return sp if split == "only" else ((out, sp) if split == "also" else out)


def attention_split(q2, q_col, q_rows, k2, k_col, v2, v_col, k_rows, B, nheads, Lq, Lk, D, scale=None,
                    key_mask=None, bias=None, out=None, split="only", hop=None, out_fmt=F16X2):
    """softmax(Q K^T * scale + bias, masked) V on split-fp16 operands (the split outputs of the projections):
    head h of Q is columns q_col + D*h of each half of q2, rows b*q_rows + i; K / V likewise with k_rows rows
    per batch element.  TMA-fed tcgen05 kernel with the probabilities in TMEM.  D in (32, 64)."""
    for t in (q2, k2, v2):
        assert isinstance(t, SplitOperand) and t.scale == 1.0
    C = nheads * D
    dev = q2.data.device
    ldo = so_ = 0
    if split != "only":
        if out is None:
            out = empty(B, Lq, C, device=dev)
        assert out.stride(2) == 1
        ldo, so_ = out.stride(1), out.stride(0)
    else:
        out = None
    sp, sp_ptr = None, None
    if split != "no":
        sp = SplitOperand(empty(B * Lq, 2 * C, dtype=torch.float16, device=dev), B * Lq, C, C, 1.0, out_fmt)
        sp_ptr = sp.data.data_ptr()
    assert out_fmt == F16X2 or (sp is not None and D == 64), "F16F8 attention output rows: head dim 64, split output"
    if scale is None:
        scale = D ** -0.5
    if key_mask is not None:
        _chk(key_mask, "key_mask", torch.uint8)
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous() and tuple(key_mask.shape) == (B, Lk)
    if bias is not None:
        _chk(bias, "bias")
        assert bias.is_contiguous() and tuple(bias.shape) == (B, nheads, Lq, Lk)
    if hop is not None:
        # the structural bias is computed inside the kernel from the hop tensor and the 5 -> 12 -> H MLP
        attn_adj, w0, b0, w1, b1 = hop
        assert bias is None and attn_adj.is_contiguous() and tuple(attn_adj.shape[1:]) == (B, Lq, Lk)
        for t_, n_ in ((attn_adj, "attn_adj"), (w0, "w0"), (b0, "b0"), (w1, "w1"), (b1, "b1")):
            _chk(t_, n_)
            assert t_.is_contiguous()
        assert w1.shape[0] == nheads
        _lib.call("ec_attention_hop_bias_next", _p(attn_adj), attn_adj.shape[0], w0.shape[0], _p(w0), _p(b0), _p(w1), _p(b1))
    if out_fmt != F16X2:
        _lib.call("ec_attention_split_fmt_next", int(out_fmt))
    _lib.call("ec_attention_tc_split", q2.data.data_ptr(), q2.rows, q2.Kp, q_col, q_rows, k2.data.data_ptr(), k2.rows,
              k2.Kp, k_col, v2.data.data_ptr(), v2.rows, v2.Kp, v_col, k_rows, _p(out), B, nheads, Lq, Lk, ldo, so_,
              float(scale), D, _p(key_mask), _p(bias), sp_ptr, C if sp is not None else 0, _stream())
    return sp if split == "only" else ((out, sp) if split == "also" else out)


def attention_split_ok(D, Lk, masked=False):
    """Can ec_attention_tc_split take this attention?  (tensor-core mode, head dim 32 / 64, key-count limits)"""
    return bool(TENSOR_CORES and ATTENTION_TC and ATTENTION_TMA and D in (32, 64)
                and Lk <= (768 if (D == 64 and not masked) else 448))


# hop-bias MLP inside the attention kernel (ec_attention_hop_bias_next).  Built, tested -- and measured
# SLOWER than the separate ec_hop_bias kernel + bias tensor: every (batch, head) CTA re-evaluates the shared 5 -> 12
# hidden layer per logit, 8 x the arithmetic of the separate kernel (head graph alone 2.20 vs 2.04 ms,
# profiles/r02_w_hop_fused_overlap.log).  Opt-in: EDGECAPE_HOP_FUSED=1.
HOP_FUSED = os.environ.get("EDGECAPE_HOP_FUSED", "0") == "1"


def hop_fused_ok(n_hops, hidden):
    """Can the attention kernel evaluate the hop-bias MLP itself (ec_attention_hop_bias_next)?"""
    return n_hops <= 8 and hidden * (n_hops + 2) + 1 <= 112


def attention_packed_split(qkv2, B, N, nheads, scale=None, out=None, split="only", key_mask=None, bias=None, hop=None,
                           out_fmt=F16X2):
    """Self-attention over a packed split-fp16 QKV operand (the split output of the QKV GEMM):
    qkv2 rows = B*N tokens, columns [q | k | v] of C = nheads*D each per half."""
    C = qkv2.K // 3
    assert qkv2.rows == B * N and qkv2.Kp == qkv2.K
    return attention_split(qkv2, 0, N, qkv2, C, qkv2, 2 * C, N, B, nheads, N, N, C // nheads, scale=scale,
                           key_mask=key_mask, bias=bias, out=out, split=split, hop=hop, out_fmt=out_fmt)


def gather_blocks(src, idx, out=None):
    """out[r] = src[idx[r]] over the leading dimension; src [n, ...] may be strided in dim 0 (inner dims contiguous),
    idx int32 [B] on the device."""
    _chk(src, "src"); _chk(idx, "idx", torch.int32)
    inner = 1
    for d in src.shape[1:]:
        inner *= d
    assert src[0].is_contiguous() and idx.is_contiguous()
    B = idx.numel()
    if out is None:
        out = empty(B, *src.shape[1:], device=src.device)
    assert out.is_contiguous()
    _lib.call("ec_gather_blocks", _p(src), src.stride(0), _p(idx), _p(out), inner, B, inner, _stream())
    return out


def hop_bias(attn_adj, w0, b0, w1, b1, out=None):
    """attn_adj [n_hops,B,K,K] -> bias [B,H,K,K] through Linear(n_hops,hidden)-ReLU-Linear(hidden,H)."""
    _chk(attn_adj, "attn_adj")
    assert attn_adj.is_contiguous()
    n_hops, B, K, _ = attn_adj.shape
    hidden, H = w0.shape[0], w1.shape[0]
    if out is None:
        out = empty(B, H, K, K, device=attn_adj.device)
    _lib.call("ec_hop_bias", _p(attn_adj), _p(w0), _p(b0), _p(w1), _p(b1), _p(out), B, K, n_hops, hidden, H,
              _stream())
    return out


def mask_accumulate_(tw, mask_s, first):
    _chk(tw, "tw"); _chk(mask_s, "mask_s")
    assert tw.is_contiguous() and mask_s.is_contiguous() and tw.numel() == mask_s.numel()
    _lib.call("ec_mask_accumulate", _p(tw), _p(mask_s), mask_s.numel(), 1 if first else 0, _stream())
    return mask_s


def kp_masks(mask_s):
    """mask_s [B,K] float -> (kp_mask, kp_mask_fixed) uint8 [B,K] (1 = padded keypoint)."""
    _chk(mask_s, "mask_s")
    B, K = mask_s.shape
    m = empty(B, K, dtype=torch.uint8, device=mask_s.device)
    mf = empty(B, K, dtype=torch.uint8, device=mask_s.device)
    _lib.call("ec_kp_masks", _p(mask_s), _p(m), _p(mf), B, K, _stream())
    return m, mf


def adj_from_edges(edges, offsets, kp_mask, K):
    """CSR edge lists (int32 device tensors) -> adj [B,2,K,K], binary [B,K,K]."""
    _chk(edges, "edges", torch.int32); _chk(offsets, "offsets", torch.int32); _chk(kp_mask, "kp_mask", torch.uint8)
    B = kp_mask.shape[0]
    adj = empty(B, 2, K, K, device=kp_mask.device)
    binary = empty(B, K, K, device=kp_mask.device)
    _lib.call("ec_adj_from_edges", _p(edges), _p(offsets), _p(kp_mask), _p(adj), _p(binary), B, K, _stream())
    return adj, binary


def soft_normalize_adj(U, kp_mask):
    _chk(U, "U"); _chk(kp_mask, "kp_mask", torch.uint8)
    B, K, _ = U.shape
    assert U.is_contiguous()
    adj = empty(B, 2, K, K, device=U.device)
    _lib.call("ec_soft_normalize_adj", _p(U), _p(kp_mask), _p(adj), B, K, _stream())
    return adj


def l2_normalize(x, eps=1e-8):
    _chk(x, "x")
    assert x.is_contiguous()
    C = x.shape[-1]
    out = empty(*x.shape, device=x.device)
    _lib.call("ec_l2_normalize", _p(x), _p(out), x.numel() // C, C, float(eps), _stream())
    return out


def edge_weights(S, binary, kp_mask, zc_w, zc_b, use_zero_conv, hops=None):
    """Gram matrix S [B,K,K] -> adj [B,2,K,K], unnorm [B,K,K]; fills hops[0], hops[1] if given."""
    _chk(S, "S"); _chk(binary, "binary"); _chk(kp_mask, "kp_mask", torch.uint8)
    B, K, _ = S.shape
    assert S.is_contiguous() and binary.is_contiguous()
    adj = empty(B, 2, K, K, device=S.device)
    unnorm = empty(B, K, K, device=S.device)
    h0 = hops[0] if hops is not None else None
    h1 = hops[1] if hops is not None and hops.shape[0] > 1 else None
    _lib.call("ec_edge_weights", _p(S), _p(binary), _p(kp_mask), float(zc_w), float(zc_b), int(use_zero_conv),
              _p(adj), _p(unnorm), _p(h0), _p(h1), B, K, _stream())
    return adj, unnorm


def markov_powers_(hops):
    """hops [H+1, B, K, K] with planes 0, 1 filled -> planes 2..H = P^h (skeleton.py:152-161), one launch."""
    _chk(hops, "hops")
    assert hops.is_contiguous() and hops.dim() == 4 and hops.shape[2] == hops.shape[3]
    H1, B, K, _ = hops.shape
    if H1 > 2:
        _lib.call("ec_markov_powers", _p(hops), H1 - 1, B, K, _stream())
    return hops


def markov_powers_ok(K):
    return ((K + 1) * (K + 1) + 4 + 32 * K) * 4 <= 227 * 1024 and K <= 256


def gcn_pack_weights(W, bias):
    """Conv1d(k=1) weight [2*dff, d(,1)] + bias [2*dff] -> packed [dff, 2d+4]."""
    _chk(W, "W"); _chk(bias, "bias")
    W = W.reshape(W.shape[0], -1).contiguous()
    dff, d = W.shape[0] // 2, W.shape[1]
    Wp = empty(dff, 2 * d + 4, device=W.device)
    _lib.call("ec_gcn_pack_weights", _p(W), _p(bias), _p(Wp), d, dff, _stream())
    return Wp


def gcn_tc_ok(B, K):
    """True when gcn() takes the tensor-core route (and can therefore hand back a split result)."""
    return TENSOR_CORES and B * K >= TC_MIN_M and ((K + 3) // 4 * 4) * 32 <= 48 * 1024


# one-kernel GCN: EDGECAPE_GCN_FUSED=2 (default) the project-first kernel (csrc/gcn_fused2_tcgen05.cu), =1 the
# aggregate-first kernel (csrc/gcn_fused_tcgen05.cu), =0 the aggregate kernel + GEMM pair
GCN_FUSED = int(os.environ.get("EDGECAPE_GCN_FUSED", "2"))


def _gcn_bias_pair(Wp, d):
    """The bias columns [2d], [2d+1] of the packed GCN weights as one contiguous [2, dff] array (cached with the split
    weights: one strided copy per checkpoint)."""
    key = (id(Wp), Wp.data_ptr(), Wp._version, tuple(Wp.shape), "gcn_bias2")
    hit = _SPLIT_WEIGHTS.get(key)
    if hit is not None and hit[0]() is Wp:
        return hit[1]
    b2 = empty(2, Wp.shape[0], device=Wp.device)
    b2.copy_(Wp[:, 2 * d:2 * d + 2].t())          # (weight repacking, once per checkpoint)
    _SPLIT_WEIGHTS[key] = (weakref.ref(Wp), b2)
    return b2


def gcn_fused_ok(B, K, d, dff):
    """Does gcn() run as ONE kernel for this shape (either formulation)?"""
    if not (GCN_FUSED and gcn_tc_ok(B, K)):
        return False
    lib = _lib.load()
    return bool((GCN_FUSED >= 2 and lib.ec_gcn_fused2_slice(K, d, dff) > 0) or lib.ec_gcn_fused_slice(K, d, dff) > 0)


def gcn(x, adj, Wp, out=None, split="no"):
    """x [B,K,d], adj [B,2,K,K] (plane 0 diagonal), packed weights -> relu(GCN) [B,K,dff].
    Tensor-core mode: one fused kernel (or, outside its shape gate, fused aggregate -> split-fp16 Z, then the tcgen05
    GEMM with a ReLU epilogue); split="only" returns the SplitOperand of the result (the A operand of the ffn2 GEMM
    that follows)."""
    _chk(x, "x"); _chk(adj, "adj"); _chk(Wp, "Wp")
    assert x.is_contiguous() and adj.is_contiguous()
    B, K, d = x.shape
    dff = Wp.shape[0]
    assert Wp.shape[1] == 2 * d + 4
    if gcn_fused_ok(B, K, d, dff):
        v2 = GCN_FUSED >= 2 and _lib.load().ec_gcn_fused2_slice(K, d, dff) > 0
        w2 = split_weight(Wp, F16F8_I32 if v2 else F16X2)
        so, so_ptr = None, None
        if split != "no":
            so = SplitOperand(empty(B * K, 2 * dff, dtype=torch.float16, device=x.device), B * K, dff, dff, 1.0)
            so_ptr = so.data.data_ptr()
        if split == "only":
            out = None
        elif out is None:
            out = empty(B, K, dff, device=x.device)
        assert out is None or out.is_contiguous()
        if v2:
            _lib.call("ec_gcn_fused2", _p(x), _p(adj), _p(_gcn_bias_pair(Wp, d)), w2.data.data_ptr(), w2.Kp, float(w2.scale),
                      _p(out), so_ptr, dff, B, K, d, dff, _stream())
        else:
            _lib.call("ec_gcn_fused", _p(x), _p(adj), _p(Wp), w2.data.data_ptr(), w2.Kp, float(w2.scale), _p(out), so_ptr,
                      dff, B, K, d, dff, _stream())
        return so if split == "only" else ((out, so) if split == "also" else out)
    if gcn_tc_ok(B, K):
        Kp = _kp(2 * d + 4)
        z2 = SplitOperand(empty(B * K, 2 * Kp, dtype=torch.float16, device=x.device), B * K, 2 * d + 4, Kp, 1.0)
        _lib.call("ec_gcn_aggregate_split", _p(x), _p(adj), _p(z2.data), B, K, d, Kp, _stream())
        if split == "only":
            return gemm_tc(z2, split_weight(Wp), act=ACT_RELU, split_out=True, fp32_out=False)[1]
        o2 = None if out is None else out.view(B * K, dff)
        res = gemm_tc(z2, split_weight(Wp), out=o2, act=ACT_RELU, split_out=(split == "also"))
        y = res[0] if split == "also" else res
        y = out if out is not None else y.view(B, K, dff)
        return (y, res[1]) if split == "also" else y
    assert split == "no", "split outputs need the tensor-core path"
    if out is None:
        out = empty(B, K, dff, device=x.device)
    assert out.is_contiguous()
    nbytes = _lib.load().ec_workspace_bytes_gcn(B, K, d, dff)
    ws = empty(nbytes // 4, device=x.device)
    _lib.call("ec_gcn", _p(x), _p(adj), _p(Wp), _p(out), B, K, d, dff, _p(ws), nbytes, _stream())
    return out


def support_weights(target, rowscale, h, w, out=None):
    """target [B,K,hm,hm] heat-maps -> pooling weights [B,K,h*w] (see include/edgecape_b200.h)."""
    _chk(target, "target"); _chk(rowscale, "rowscale")
    assert target.is_contiguous()
    B, K, hm_h, hm_w = target.shape
    if out is None:
        out = empty(B, K, h * w, device=target.device)
    assert out.stride(2) == 1 and out.stride(0) == K * out.stride(1)
    _lib.call("ec_support_weights", _p(target), _p(rowscale), _p(out), out.stride(1), B * K, hm_h, hm_w, h, w,
              _stream())
    return out


def sine_pe_coords(coord, num_feats=128, temperature=10000.0, scale=6.283185307179586, out=None):
    """coord [..., 2] (x,y) -> [..., 2*num_feats] DETR sine encoding, channels [y-half | x-half]."""
    _chk(coord, "coord")
    assert coord.is_contiguous() and coord.shape[-1] == 2
    M = coord.numel() // 2
    if out is None:
        out = empty(*coord.shape[:-1], 2 * num_feats, device=coord.device)
    o2 = out.reshape(M, -1) if out.is_contiguous() else out
    assert o2.dim() == 2 and o2.stride(1) == 1
    _lib.call("ec_sine_pe_coords", _p(coord), _p(o2), o2.stride(0), M, num_feats, float(temperature), float(scale),
              _stream())
    return out


def proposal(sim, h, w):
    """sim [B,K,h*w] -> (proposal_for_loss [B,K,2], proposals [B,K,2], argmax int64 [B,K])."""
    _chk(sim, "sim")
    assert sim.is_contiguous()
    B, K = sim.shape[:2]
    pl = empty(B, K, 2, device=sim.device)
    pr = empty(B, K, 2, device=sim.device)
    am = empty(B, K, dtype=torch.int64, device=sim.device)
    _lib.call("ec_proposal", _p(sim), _p(pl), _p(pr), _p(am), B * K, h, w, _stream())
    return pl, pr, am


def point_update(bi, delta, out=None):
    """sigmoid(inverse_sigmoid(bi) + delta); bi [...,2] contiguous, delta [M,2] view."""
    _chk(bi, "bi"); _chk(delta, "delta")
    assert bi.is_contiguous()
    M = bi.numel() // 2
    d2 = delta.reshape(M, 2) if delta.is_contiguous() else delta
    assert d2.stride(1) == 1
    if out is None:
        out = empty(*bi.shape, device=bi.device)
    _lib.call("ec_point_update", _p(bi), _p(d2), d2.stride(0), _p(out), M, _stream())
    return out


def decode_preds(points, center_scale, W, H, use_udp=False, out=None):
    """points [B,K,2] in [0,1], center_scale [B,4] -> preds [B,K,3] in image coordinates (TwoStageHead.decode)."""
    _chk(points, "points"); _chk(center_scale, "center_scale")
    assert points.is_contiguous() and center_scale.is_contiguous()
    B, K, _ = points.shape
    if out is None:
        out = empty(B, K, 3, device=points.device)
    _lib.call("ec_decode_preds", _p(points), _p(center_scale), _p(out), B, K, float(W), float(H), int(use_udp), _stream())
    return out


def im2col_patches(img, P, ldc=None):
    _chk(img, "img")
    assert img.is_contiguous()
    B, C3, H, W = img.shape
    assert C3 == 3
    h0, w0 = H // P, W // P
    ldc = ldc or 3 * P * P
    cols = empty(B * h0 * w0, ldc, device=img.device)
    _lib.call("ec_im2col_patches", _p(img), _p(cols), B, H, W, P, ldc, None, 0, _stream())
    return cols


def im2col_patches_split(images, P):
    """list of [b,3,H,W] images -> SplitOperand [sum(b)*h0*w0, 2*Kp] of the patch matrix (Kp = 3*P*P rounded to 64)."""
    H, W = images[0].shape[-2:]
    h0, w0 = H // P, W // P
    S = h0 * w0
    K = 3 * P * P
    Kp = _kp(K)
    rows = sum(int(im.shape[0]) for im in images) * S
    so = SplitOperand(empty(rows, 2 * Kp, dtype=torch.float16, device=images[0].device), rows, K, Kp, 1.0)
    r = 0
    for im in images:
        _chk(im, "img")
        assert im.is_contiguous() and tuple(im.shape[-2:]) == (H, W)
        b = int(im.shape[0])
        _lib.call("ec_im2col_patches", _p(im), None, b, H, W, P, Kp, so.data[r * S:].data_ptr(), Kp, _stream())
        r += b
    return so


def interp_pos_embed(pos_embed, h0, w0, offset=0.1):
    """pos_embed [1, 1+M*M, C] -> [1+h0*w0, C] (DINOv2 interpolate_pos_encoding)."""
    _chk(pos_embed, "pos_embed")
    pe = pos_embed.reshape(-1, pos_embed.shape[-1]).contiguous()
    N = pe.shape[0] - 1
    Mg = int(round(N ** 0.5))
    assert Mg * Mg == N
    out = empty(1 + h0 * w0, pe.shape[1], device=pe.device)
    _lib.call("ec_interp_pos_embed", _p(pe), _p(out), Mg, h0, w0, pe.shape[1], float(offset), _stream())
    return out


def write_cls_(tokens, cls, pos0):
    _chk(tokens, "tokens")
    B, T, C = tokens.shape
    assert tokens.stride(2) == 1
    _lib.call("ec_write_cls", _p(cls), _p(pos0), _p(tokens), B, tokens.stride(0), C, _stream())
    return tokens


def metrics_accumulate_(counters, pred, gt, valid, norm, thr, auc_steps=20):
    """counters (fp64, >= T+4) += [per-sample PCK@thr..., NME, AUC, EPE, 1] for every sample of the batch."""
    _chk(pred, "pred"); _chk(gt, "gt"); _chk(valid, "valid", torch.uint8); _chk(norm, "norm"); _chk(thr, "thr")
    _chk(counters, "counters", torch.float64)
    B, K, _ = pred.shape
    assert pred.is_contiguous() and gt.is_contiguous() and valid.is_contiguous() and norm.is_contiguous()
    assert counters.numel() >= thr.numel() + 4
    _lib.call("ec_metrics_accumulate", _p(pred), _p(gt), _p(valid), _p(norm), _p(thr), thr.numel(), int(auc_steps),
              _p(counters), B, K, _stream())
    return counters


def pck_accumulate_(counters, pred, gt, valid, norm, thr):
    _chk(pred, "pred"); _chk(gt, "gt"); _chk(valid, "valid", torch.uint8); _chk(norm, "norm"); _chk(thr, "thr")
    _chk(counters, "counters", torch.float64)
    B, K, _ = pred.shape
    assert pred.is_contiguous() and gt.is_contiguous() and valid.is_contiguous() and norm.is_contiguous()
    assert counters.numel() >= thr.numel() + 1
    _lib.call("ec_pck_accumulate", _p(pred), _p(gt), _p(valid), _p(norm), _p(thr), thr.numel(), _p(counters), B, K,
              _stream())
    return counters
