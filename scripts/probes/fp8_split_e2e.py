"""Offline (CPU) end-to-end accuracy probe for the next GEMM step.

Runs the oracle's detector forward (configs[1] at reduced batch: ViT-B/14, 256x256, K = 100) with every nn.Linear
on the path replaced by an emulation of a tensor-core scheme, and reports how far the ViT features, the similarity
heat-map, its arg-max and the final keypoints move from the fp32 run:

    3xfp16      a_lo.b_hi + a_hi.b_lo + a_hi.b_hi on fp16 operands (what gemm_tcgen05.cu does today)
    fp16+2xfp8  a_hi.b_hi on fp16; both cross terms with e4m3 operands (power-of-two scale per tensor) -- 2 units of
                tensor time instead of 3
    fp8 static  the same two cross terms with STATIC power-of-two scales, chosen so that all three products carry the
                weight scale s_w of ops.split_weight and can share ONE TMEM accumulator, with no absmax pass over the
                activations:  a_lo8 = e4m3(a_lo 2^11), b_hi8 = e4m3(b_hi s_w 2^-11), a_hi8 = e4m3(a_hi), b_lo8 = e4m3(b_lo s_w)
    fp16+1xfp8  a_hi.b_hi on fp16; only a_lo.b_hi8 kept (weights' low part dropped) -- 1.5 units
    fp16        a_hi.b_hi only (plain fp16 GEMM) -- 1 unit

Test infrastructure only (imports oracle/); nothing here is on the product path.  Products are formed in fp64 from the
quantised operands, i.e. the fp32 accumulation of the tensor core is not modelled (it adds ~1e-6).
usage: python scripts/probes/fp8_split_e2e.py [case]      (default c2_vitb_256_k100)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from edgecape_b200.config import state_dict_shapes  # noqa: E402
from edgecape_b200.synthetic import make_state_dict  # noqa: E402
from oracle import edgecape_oracle  # noqa: E402
from oracle.gen_golden import build_case  # noqa: E402

_real_linear = F.linear
MIN_ROWS, MIN_N, MIN_K = 64, 32, 32      # ops.TC_MIN_*: smaller problems run on the exact fp32 kernel


def split16(a):
    hi = a.to(torch.float16).to(torch.float64)
    lo = (a.to(torch.float64) - hi).to(torch.float16).to(torch.float64)
    return hi, lo


def q8(a):
    amax = a.abs().max().item()
    if amax == 0.0 or not np.isfinite(amax):
        return a
    s = 2.0 ** np.floor(np.log2(240.0 / amax))
    return (a * s).to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64) / s


def e4m3(a):
    return a.to(torch.float32).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float64)


def weight_scale(w):
    """ops.split_weight: power of two that brings the tensor's absmax to ~16384."""
    amax = float(w.abs().max())
    if amax <= 0 or not np.isfinite(amax):
        return 1.0
    return 2.0 ** max(-8, min(14, int(np.floor(np.log2(16384.0 / amax)))))


def make_linear(scheme):
    def linear(x, w, b=None):
        rows = x.numel() // x.shape[-1]
        if scheme == "fp32" or rows < MIN_ROWS or w.shape[0] < MIN_N or w.shape[1] < MIN_K:
            return _real_linear(x, w, b)
        x2 = x.reshape(rows, x.shape[-1])
        xh, xl = split16(x2)
        wh, wl = split16(w)
        y = xh @ wh.T
        if scheme == "3xfp16":
            y = y + xl @ wh.T + xh @ wl.T
        elif scheme == "fp16+2xfp8":
            y = y + q8(xl) @ q8(wh).T + q8(xh) @ q8(wl).T
        elif scheme == "fp8 static":
            sw = weight_scale(w)
            cross = e4m3(xl * 2.0 ** 11) @ e4m3(wh * (sw * 2.0 ** -11)).T + e4m3(xh) @ e4m3(wl * sw).T
            y = y + cross / sw
        elif scheme == "fp16+1xfp8":
            y = y + q8(xl) @ q8(wh).T
        elif scheme != "fp16":
            raise ValueError(scheme)
        y = y.to(x.dtype)
        if b is not None:
            y = y + b
        return y.reshape(*x.shape[:-1], w.shape[0])
    return linear


def run(sd, cfg, data, scheme):
    F.linear = make_linear(scheme)
    try:
        with torch.no_grad():
            return edgecape_oracle.detector_forward_test(sd, cfg, data, torch.float32)
    finally:
        F.linear = _real_linear


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def main(case="c2_vitb_256_k100"):
    torch.manual_seed(0)
    cfg, data, wseed = build_case(case)
    sd = make_state_dict(state_dict_shapes(cfg), wseed)
    ref = run(sd, cfg, data, "fp32")
    am_ref = ref["similarity_map"].flatten(2).argmax(-1)
    print(f"case {case}: relative max error against the fp32 run (parity bar 1e-3; arg-max must not move)")
    print(f"{'scheme':12s} {'ViT features':>13s} {'heat-map':>10s} {'arg-max moved':>14s} {'points':>10s} {'preds':>10s}")
    for scheme in ("3xfp16", "fp16+2xfp8", "fp8 static", "fp16+1xfp8", "fp16"):
        got = run(sd, cfg, data, scheme)
        am = got["similarity_map"].flatten(2).argmax(-1)
        print(f"{scheme:12s} {rel(got['feature_q'], ref['feature_q']):13.2e} {rel(got['similarity_map'], ref['similarity_map']):10.2e} "
              f"{int((am != am_ref).sum()):8d}/{am.numel():<5d} {rel(got['points'], ref['points']):10.2e} {rel(got['preds'], ref['preds']):10.2e}")


if __name__ == "__main__":
    main(*sys.argv[1:2])
