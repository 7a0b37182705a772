"""GPU: every C-ABI kernel against the CPU oracle (oracle/edgecape_oracle.py) or the plain torch
fp32 statement of the same op, on seeded inputs including ragged / unaligned / masked cases.
Tolerances are relative to the tensor's max magnitude; indices must be bit-exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from edgecape_b200 import ops  # noqa: E402
from oracle import edgecape_oracle as O  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def close(got, want, tol=2e-5, what=""):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs().max().item() / (want.abs().max().item() + 1e-12)
    assert err < tol, f"{what}: rel err {err:.3e} >= {tol}"
    return err


# ----------------------------------------------------------------------------------- gemm
@pytest.mark.parametrize("M,N,K", [(1600, 256, 768), (100, 2, 256), (37, 53, 19), (324, 768, 588), (128, 128, 16),
                                   (5, 7, 3), (1, 1, 1)])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_RELU, ops.ACT_GELU, ops.ACT_TANH])
def test_gemm_linear_epilogues(M, N, K, act):
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    g, r = rnd(N, seed=4), rnd(M, N, seed=5)
    y = F.linear(x, w, b)
    y = [y, F.relu(y), F.gelu(y), torch.tanh(y)][act]
    want = r + g * y
    got = ops.linear(x.to(dev()), w.to(dev()), b.to(dev()), act=act, colscale=g.to(dev()), residual=r.to(dev()))
    close(got, want, what=f"linear {M}x{N}x{K} act{act}")
    want = (y + 1) * r
    got = ops.linear(x.to(dev()), w.to(dev()), b.to(dev()), act=act, residual=r.to(dev()), res_mode=ops.RES_GATE)
    close(got, want, what="gate")


def test_gemm_inplace_residual_and_strided_views():
    B, S, C, N = 3, 36, 64, 48
    t = rnd(B, S + 1, N, seed=1).to(dev())
    t0 = t.clone()
    a = rnd(B, S, C, seed=2).to(dev())
    w = rnd(N, C, seed=3, scale=0.1).to(dev())
    pos = rnd(S, N, seed=4).to(dev())
    # out is a strided 3-D view (cls row skipped), residual broadcast over the batch
    ops.gemm(a, w, out=t[:, 1:, :], residual=pos, res_mode=ops.RES_ADD)
    want = torch.einsum("bsc,nc->bsn", a.cpu(), w.cpu()) + pos.cpu()
    close(t[:, 1:, :], want, what="strided out")
    assert torch.equal(t[:, 0, :], t0[:, 0, :])
    # in-place residual (C aliases R)
    x = rnd(50, 64, seed=5).to(dev())
    acc = rnd(50, 48, seed=6).to(dev())
    want = acc.cpu() + x.cpu() @ w.cpu().T
    ops.linear(x, w, None, residual=acc, out=acc)
    close(acc, want, what="in-place")


@pytest.mark.parametrize("K", [100, 17, 200])
def test_gemm_batched_nn_and_tn(K):
    B, d = 3, 64
    A = rnd(B, K, K, seed=1).to(dev())
    X = rnd(B, K, d, seed=2).to(dev())
    close(ops.gemm(A, X, b_kmajor=False), torch.bmm(A.cpu(), X.cpu()), what="bmm NN")
    close(ops.gemm(X, X, b_kmajor=True), torch.bmm(X.cpu(), X.cpu().transpose(1, 2)), what="bmm TN")
    # column-slice views of a wider buffer
    buf = rnd(B * K, 2 * d, seed=3).to(dev())
    w = rnd(24, d, seed=4).to(dev())
    close(ops.linear(buf[:, d:], w), buf.cpu()[:, d:] @ w.cpu().T, what="ld view")


# ------------------------------------------------------------------------------- rowwise
@pytest.mark.parametrize("M,C", [(650, 768), (7, 64), (33, 1024), (5, 100), (100, 256), (13, 384), (9, 512)])
def test_layernorm(M, C):
    x, r = rnd(M, C, seed=1, scale=3.0), rnd(M, C, seed=2)
    w, b = 1 + 0.1 * rnd(C, seed=3), 0.1 * rnd(C, seed=4)
    want = F.layer_norm(x + r, (C,), w, b, 1e-6)
    s = ops.empty(M, C)
    got = ops.layernorm(x.to(dev()), w.to(dev()), b.to(dev()), 1e-6, residual=r.to(dev()), sum_out=s)
    close(got, want, what="ln")
    close(s, x + r, tol=1e-7, what="sum_out")


@pytest.mark.parametrize("C", [64, 256])
def test_layernorm_drops_cls_row_by_striding(C):
    B, S = 3, 16
    t = rnd(B, S + 1, C, seed=1)
    w, b = 1 + 0.1 * rnd(C, seed=3), 0.1 * rnd(C, seed=4)
    want = F.layer_norm(t[:, 1:], (C,), w, b, 1e-5)
    got = ops.layernorm(t.to(dev())[:, 1:, :], w.to(dev()), b.to(dev()), 1e-5)
    close(got, want, what="ln seg")


def test_add_copy_axpby_l2():
    B, T, S, C = 2, 12, 9, 32
    x, p = rnd(B, T, C, seed=1), rnd(S, C, seed=2)
    want = x.clone()
    want[:, :S] += p
    close(ops.add_rows_(x.to(dev()), p.to(dev()), S), want, tol=1e-7)
    xd = x.to(dev())
    out = torch.zeros(B, S, 2 * C, device=dev())
    ops.copy_rows(xd[:, :S, :], out[:, :, :C])
    ops.copy_rows(p.to(dev()), out.view(B * S, 2 * C)[:, C:], bcast_rows=S)
    close(out, torch.cat((x[:, :S], p[None].expand(B, -1, -1)), dim=-1), tol=1e-7)
    close(ops.axpby(xd, xd, 2.0, 1.0, 5.0), (2 * x + x) / 5.0, tol=1e-6)
    close(ops.l2_normalize(xd), x / (x.norm(dim=-1, keepdim=True) + 1e-8), tol=1e-6)


# ----------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,H,Lq,Lk,D", [(2, 12, 325, 325, 64), (3, 8, 100, 324, 64), (2, 8, 424, 424, 32),
                                         (2, 4, 17, 17, 16), (1, 8, 324, 100, 64), (2, 8, 200, 200, 32)])
def test_attention_plain_and_masked(B, H, Lq, Lk, D):
    E = H * D
    q, k, v = rnd(B, Lq, E, seed=1), rnd(B, Lk, E, seed=2), rnd(B, Lk, E, seed=3)
    mask = torch.zeros(B, Lk, dtype=torch.bool)
    mask[:, Lk - Lk // 3:] = True
    mask[0, 0] = True
    for m in (None, mask):
        qh = q.view(B, Lq, H, D).transpose(1, 2) * D ** -0.5
        kh = k.view(B, Lk, H, D).transpose(1, 2)
        vh = v.view(B, Lk, H, D).transpose(1, 2)
        s = qh @ kh.transpose(-1, -2)
        if m is not None:
            s = s.masked_fill(m[:, None, None, :], float("-inf"))
        want = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E)
        got = ops.attention(q.to(dev()), k.to(dev()), v.to(dev()), H,
                            key_mask=None if m is None else m.to(torch.uint8).to(dev()))
        close(got, want, what=f"attn {B},{H},{Lq},{Lk},{D} mask={m is not None}")


def test_attention_packed_qkv_views_and_bias_matches_oracle_mha():
    """BiasedMultiheadAttention (utils/bias_attn.py:106-231) via oracle.mha with attn_bias."""
    B, K, d, H = 2, 100, 256, 8
    x = rnd(B, K, d, seed=1)
    wq, wk, wv, wo = (rnd(d, d, seed=s, scale=d ** -0.5) for s in (2, 3, 4, 5))
    bq, bk, bv, bo = (0.1 * rnd(d, seed=s) for s in (6, 7, 8, 9))
    hops = torch.rand(5, B, K, K, generator=torch.Generator().manual_seed(10))
    w0, b0, w1, b1 = rnd(12, 5, seed=11), rnd(12, seed=12), rnd(8, 12, seed=13), rnd(8, seed=14)
    mask = torch.zeros(B, K, dtype=torch.bool)
    mask[1, 70:] = True
    bias = F.linear(F.relu(F.linear(hops.permute(1, 2, 3, 0), w0, b0)), w1, b1).permute(0, 3, 1, 2)
    want = O.mha(x, x, x, wq, wk, wv, bq, bk, bv, wo, bo, H, mask, attn_bias=bias)
    D = dev()
    qkv = ops.linear(x.to(D).view(B * K, d), torch.cat((wq, wk, wv)).to(D), torch.cat((bq, bk, bv)).to(D)).view(B, K, 3 * d)
    gb = ops.hop_bias(hops.to(D), w0.to(D), b0.to(D), w1.to(D), b1.to(D))
    close(gb, bias, what="hop bias")
    a = ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], H, key_mask=mask.to(torch.uint8).to(D), bias=gb)
    got = ops.linear(a.view(B * K, d), wo.to(D), bo.to(D)).view(B, K, d)
    close(got, want, what="biased mha")


# -------------------------------------------------------------------------------- graph
def _edges(skeleton, D):
    from edgecape_b200.skeleton import edges_to_csr
    return edges_to_csr(skeleton, D)


@pytest.mark.parametrize("K", [100, 17, 5])
def test_adjacency_from_edges_and_soft_normalize(K):
    rng = np.random.default_rng(K)
    B = 4
    from edgecape_b200.synthetic import random_skeleton
    skel = [random_skeleton(rng, K, K - K // 4), [], random_skeleton(rng, K, K, "chain"), [[0, 1], [1, 0], [0, 1]]]
    mask = torch.zeros(B, K, dtype=torch.bool)
    mask[0, K - K // 4:] = True
    mask[1, :] = True
    mask[3, 1] = True
    want = O.adj_from_edges(skel, K, mask, torch.float32)
    D = dev()
    e, o = _edges(skel, D)
    adj, binary = ops.adj_from_edges(e, o, mask.to(torch.uint8).to(D), K)
    close(adj, want, tol=1e-6, what="adj")
    assert torch.equal(binary.cpu() > 0, want[:, 1] > 0)
    U = torch.rand(B, K, K, generator=torch.Generator().manual_seed(1))
    close(ops.soft_normalize_adj(U.to(D), mask.to(torch.uint8).to(D)), O.soft_normalize_adj(U, mask), tol=1e-6)


@pytest.mark.parametrize("B,K,d,dff", [(4, 100, 256, 384), (2, 17, 256, 64), (2, 200, 256, 384), (3, 100, 256, 768)])
@pytest.mark.parametrize("tensor_cores", [True, False, "fused", "fused1"], ids=["tcgen05", "simt", "fused", "fused_aggregate_first"])
def test_gcn_matches_oracle(B, K, d, dff, tensor_cores, monkeypatch):
    fused = {"fused": 2, "fused1": 1}.get(tensor_cores, 0)
    tensor_cores = bool(tensor_cores)
    monkeypatch.setattr(ops, "TENSOR_CORES", tensor_cores)
    monkeypatch.setattr(ops, "GCN_FUSED", fused)
    if fused and not ops.gcn_fused_ok(B, K, d, dff):
        pytest.skip("shape outside the fused kernel (falls back to the two-kernel path)")
    x = rnd(B, K, d, seed=1)
    W, b = rnd(2 * dff, d, 1, seed=2, scale=d ** -0.5), 0.1 * rnd(2 * dff, seed=3)
    mask = torch.zeros(B, K, dtype=torch.bool)
    mask[0, K - K // 5:] = True
    U = torch.rand(B, K, K, generator=torch.Generator().manual_seed(4))
    adj = O.soft_normalize_adj(U, mask)
    want = O.gcn(x, adj, W, b)
    D = dev()
    Wp = ops.gcn_pack_weights(W.to(D), b.to(D))
    tol = 5e-5 if fused == 2 else 2e-5      # the project-first kernel runs its cross terms on e4m3
    got = ops.gcn(x.to(D), adj.to(D).contiguous(), Wp)
    close(got, want, tol=tol, what="gcn")
    if tensor_cores and B * K >= ops.TC_MIN_M:
        so = ops.gcn(x.to(D), adj.to(D).contiguous(), Wp, split="only")
        close(so.data[:, :dff].float() + so.data[:, so.Kp:so.Kp + dff].float(), want.reshape(B * K, dff), tol=tol,
              what="gcn split")


@pytest.mark.parametrize("B,K,d,dff", [(64, 100, 256, 384), (5, 97, 128, 192), (3, 112, 256, 768), (2, 64, 64, 64),
                                      (3, 128, 256, 384), (4, 16, 256, 384), (70, 1, 64, 128), (170, 100, 256, 384),
                                      (101, 50, 256, 576), (4, 200, 256, 384), (3, 136, 256, 768), (80, 200, 256, 384),
                                      (2, 256, 256, 384), (3, 200, 128, 128)])
@pytest.mark.parametrize("variant", [2, 1], ids=["project_first", "aggregate_first"])
def test_gcn_fused_kernel(B, K, d, dff, variant, monkeypatch):
    """One-kernel GCN (gcn_fused2_tcgen05.cu: project first, e4m3 cross terms, A1 in tensor memory, persistent CTAs --
    (170, ...) and (101, ...) give every CTA several items, with two and with three channel slices; K > 128 (configs[4]:
    K = 200) runs as clusters of two CTAs that exchange their T1 rows through distributed shared memory;
    gcn_fused_tcgen05.cu: aggregate first, K <= 128) vs the fp64 oracle and vs the two-kernel tensor-core path: general
    (non 0/1) diagonal plane, masked rows, fp32 and split outputs."""
    if variant == 1 and K > 128:
        pytest.skip("the aggregate-first kernel takes K <= 128")
    monkeypatch.setattr(ops, "TENSOR_CORES", True)
    monkeypatch.setattr(ops, "GCN_FUSED", variant)
    assert ops.gcn_fused_ok(B, K, d, dff)
    lib = ops._lib.load()
    assert (lib.ec_gcn_fused2_slice if variant == 2 else lib.ec_gcn_fused_slice)(K, d, dff) > 0
    tol = 5e-5 if variant == 2 else 2e-5          # e4m3 cross terms: ~1e-5 of max|Y| (three fp16 products: 3e-6)
    x = rnd(B, K, d, seed=11)
    W, b = rnd(2 * dff, d, 1, seed=12, scale=d ** -0.5), 0.1 * rnd(2 * dff, seed=13)
    mask = torch.zeros(B, K, dtype=torch.bool)
    mask[0, K - K // 5:] = True
    mask[B - 1, ::3] = True
    U = torch.rand(B, K, K, generator=torch.Generator().manual_seed(14))
    adj = O.soft_normalize_adj(U, mask)
    if B > 2:   # a diagonal plane that is not 0/1 exercises the in-place rescale of the X tiles
        adj[1, 0] = torch.diag_embed(torch.rand(K, generator=torch.Generator().manual_seed(15)) * 2 - 0.5)
    want = O.gcn(x.double(), adj.double(), W.double(), b.double()).float()
    D = dev()
    Wp = ops.gcn_pack_weights(W.to(D), b.to(D))
    xd, ad = x.to(D), adj.to(D).contiguous()
    ops.gcn(xd, ad, Wp)                            # (the first call also splits the weights)
    launches = ops._lib.launch_count()
    got = ops.gcn(xd, ad, Wp)
    assert ops._lib.launch_count() - launches == 1
    close(got, want, tol=tol, what="gcn fused")
    y2, so = ops.gcn(xd, ad, Wp, split="also")
    assert torch.equal(y2, got)
    close(so.data[:, :dff].float() + so.data[:, so.Kp:so.Kp + dff].float(), want.reshape(B * K, dff), tol=tol,
          what="gcn fused split")
    so2 = ops.gcn(xd, ad, Wp, split="only")
    assert torch.equal(so2.data, so.data)
    monkeypatch.setattr(ops, "GCN_FUSED", 0)
    close(got, ops.gcn(xd, ad, Wp), tol=tol, what="fused vs two-kernel")


def test_edge_weights_and_markov_match_oracle():
    B, K, d = 3, 100, 256
    f = rnd(B, K, d, seed=1)
    mask = torch.zeros(B, K, dtype=torch.bool)
    mask[0, 80:] = True
    mask[2, :] = True
    binary = (torch.rand(B, K, K, generator=torch.Generator().manual_seed(2)) > 0.9)
    binary = (binary | binary.transpose(1, 2)) & ~mask[:, :, None] & ~mask[:, None, :]
    zw, zb = 0.6, 0.05
    fn = f / (f.norm(dim=-1, keepdim=True) + 1e-8)
    S = fn @ fn.transpose(1, 2)
    S = (S + S.transpose(1, 2)) / 2 * zw + zb
    U = F.relu(binary.float() + S)
    want_adj = O.soft_normalize_adj(U, mask)
    P = want_adj[:, 1] / (want_adj[:, 1].sum(-1, keepdim=True) + 1e-8)
    want_h = torch.stack([torch.matrix_power(P, h) for h in range(5)])
    D = dev()
    g = ops.gemm(ops.l2_normalize(f.to(D)), ops.l2_normalize(f.to(D)), b_kmajor=True)
    hops = ops.empty(5, B, K, K)
    adj, un = ops.edge_weights(g, binary.float().to(D), mask.to(torch.uint8).to(D), zw, zb, True, hops)
    for h in range(2, 5):
        ops.gemm(hops[h // 2], hops[h - h // 2], out=hops[h], b_kmajor=False)
    close(adj, want_adj, what="edge adj")
    close(un, U * (~mask[:, :, None]) * (~mask[:, None, :]), what="unnorm")
    close(hops, want_h, what="markov")


# ---------------------------------------------------------------------------- head ops
@pytest.mark.parametrize("h,hm", [(18, 64), (4, 64), (27, 64), (16, 64)])
def test_support_pooling_matches_reference_form(h, hm):
    B, K, C = 2, 9, 48
    feat = rnd(B, h * h, C, seed=1)
    g = torch.Generator().manual_seed(2)
    target = torch.rand(B, K, hm, hm, generator=g) * (torch.rand(B, K, hm, hm, generator=g) > 0.9)
    target[0, 0] = 0            # empty heat-map: sum + 1e-8 guard
    mask = torch.ones(B, K)
    mask[1, 5:] = 0
    f4 = feat.transpose(1, 2).reshape(B, C, h, h)
    up = F.interpolate(f4, size=(hm, hm), mode="bilinear", align_corners=False)
    t = target / (target.sum(-1).sum(-1)[:, :, None, None] + 1e-8)
    want = (t.flatten(2) @ up.flatten(2).permute(0, 2, 1)) * mask[..., None]
    D = dev()
    tw = ops.support_weights(target.to(D), mask.to(D), h, h)
    got = ops.gemm(tw, feat.to(D), b_kmajor=False)
    close(got, want, tol=5e-5, what="pool")


def test_sine_pe_matches_oracle():
    c = torch.rand(3, 50, 2, generator=torch.Generator().manual_seed(1))
    close(ops.sine_pe_coords(c.to(dev())), O.sine_pe_coords(c), tol=5e-6, what="pe coords")
    from edgecape_b200.positional_encoding import SinePositionalEncoding
    pe = SinePositionalEncoding(num_feats=128, normalize=True)
    for h in (18, 4, 27):
        close(pe.grid_tokens(h, h, dev()), O.sine_pe_grid(h, h, torch.float32), tol=5e-6, what="pe grid")


@pytest.mark.parametrize("h", [18, 4, 27, 16])
def test_proposal_argmax_bit_exact(h):
    B, K, S = 2, 100, h * h
    sim = rnd(B, K, S, seed=h, scale=3.0)
    sim[0, 0, 5] = sim[0, 0].max() + 1          # corner / border maxima
    sim[0, 1, S - 1] = sim[0, 1].max() + 1
    sim[0, 2, :] = 0.25                           # full tie -> first index
    sim[0, 3, 7] = sim[0, 3, 3] = sim[0, 3].max() + 2   # two-way tie -> first
    sm = sim.softmax(-1)
    gy, gx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, h - 0.5, h), indexing="ij")
    grid = torch.stack((gx, gy), -1).reshape(S, 2)
    want_pl = (sm[..., None] * grid).sum(2) / h
    am = sim.argmax(-1)
    local = F.max_pool2d(F.one_hot(am, S).reshape(B, K, h, h).float(), 3, 1, 1).reshape(B, K, S)
    l = sm * local
    l = l / (l.sum(-1, keepdim=True) + 1e-10)
    want_pr = (l[..., None] * grid).sum(2) / h
    pl, pr, gam = ops.proposal(sim.to(dev()), h, h)
    assert torch.equal(gam.cpu(), am)
    close(pl, want_pl, tol=1e-5, what="prop loss")
    close(pr, want_pr, tol=1e-5, what="prop")


def test_point_update_and_masks():
    bi = torch.rand(4, 100, 2, generator=torch.Generator().manual_seed(1))
    bi[0, 0] = torch.tensor([0.0, 1.0])
    bi[0, 1] = torch.tensor([1e-4, 0.9999])
    delta = rnd(400, 2, seed=2)
    want = (O.inverse_sigmoid(bi) + delta.view(4, 100, 2)).sigmoid()
    close(ops.point_update(bi.to(dev()), delta.to(dev())), want, tol=2e-6)
    tw = torch.ones(3, 7, 1)
    tw[0, 4:] = 0
    tw[2, :] = 0
    D = dev()
    ms = ops.empty(3, 7)
    ops.mask_accumulate_(tw.to(D).reshape(3, 7), ms, True)
    ops.mask_accumulate_(tw.to(D).reshape(3, 7), ms, False)
    m, mf = ops.kp_masks(ms)
    want_m = ~(tw.squeeze(-1).bool())
    assert torch.equal(m.cpu().bool(), want_m)
    wf = want_m.clone()
    wf[2, 0] = False
    assert torch.equal(mf.cpu().bool(), wf)


# ----------------------------------------------------------------------------- ViT ops
def test_patch_embed_pos_embed_and_vit_block_match_oracle():
    from oracle.dinov2_oracle import interpolate_pos_embed
    B, P, R, C = 2, 14, 256, 64
    img = rnd(B, 3, R, R, seed=1)
    w, b = rnd(C, 3, P, P, seed=2, scale=0.05), rnd(C, seed=3)
    want = F.conv2d(img, w, b, stride=P).flatten(2).transpose(1, 2)
    D = dev()
    cols = ops.im2col_patches(img.to(D), P)
    got = ops.linear(cols, w.reshape(C, -1).to(D), b.to(D)).view(B, -1, C)
    close(got, want, what="patch embed")
    pos = rnd(1, 1 + 37 * 37, C, seed=4)
    for h0 in (18, 16, 27, 37):
        close(ops.interp_pos_embed(pos.to(D), h0, h0, 0.1), interpolate_pos_embed(pos, h0, h0, 0.1)[0], tol=5e-6,
              what=f"pos {h0}")
    pos5 = rnd(1, 26, C, seed=5)
    for h0 in (4, 6):
        close(ops.interp_pos_embed(pos5.to(D), h0, h0, 0.1), interpolate_pos_embed(pos5, h0, h0, 0.1)[0], tol=5e-6)


def test_pck_counters():
    B, K = 5, 20
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(B, K, 2, generator=g) * 200
    pred = gt + torch.randn(B, K, 2, generator=g) * 20
    valid = torch.rand(B, K, generator=g) > 0.3
    valid[4] = False
    norm = torch.full((B, 2), 200.0)
    thr = torch.tensor([0.05, 0.1, 0.15, 0.2, 0.25])
    want = torch.zeros(6, dtype=torch.float64)
    for b in range(B):
        d = ((pred[b] - gt[b]) / norm[b]).norm(dim=-1)
        for t in range(5):
            if valid[b].any():
                want[t] += (d[valid[b]] < thr[t]).double().mean()
        want[5] += 1
    D = dev()
    c = torch.zeros(6, dtype=torch.float64, device=D)
    ops.pck_accumulate_(c, pred.to(D), gt.to(D), valid.to(torch.uint8).to(D), norm.to(D), thr.to(D))
    assert torch.allclose(c.cpu(), want, atol=1e-9)


def test_metric_counters_match_report_metric_oracle():
    """ec_metrics_accumulate vs the restated mmpose metrics + the reference's per-sample reduction, including a sample
    without valid keypoints, a zero normaliser and a negative one."""
    from oracle import metrics_oracle as mo
    from edgecape_b200.parallel import PCK_THRESHOLDS, new_metric_counters, summarize_metrics
    B, K = 7, 33
    g = torch.Generator().manual_seed(1)
    gt = torch.rand(B, K, 2, generator=g) * 200
    pred = gt + torch.randn(B, K, 2, generator=g) * 25
    valid = torch.rand(B, K, generator=g) > 0.3
    valid[4] = False
    bbox = torch.tensor([200.0, 150.0, 80.0, 300.0, 120.0, 0.0, -5.0])
    norm = torch.stack((bbox, bbox), dim=1)
    want = mo.report_metric(list(pred.double().numpy()), list(gt.double().numpy()), list(valid.numpy()),
                            list(bbox.double().numpy()), PCK_THRESHOLDS)
    D = dev()
    c = new_metric_counters(D)
    thr = torch.tensor(PCK_THRESHOLDS, dtype=torch.float32)
    for lo in (0, 4):                       # two calls accumulate into the same vector
        hi = 4 if lo == 0 else B
        ops.metrics_accumulate_(c, pred[lo:hi].contiguous().to(D), gt[lo:hi].contiguous().to(D),
                                valid[lo:hi].to(torch.uint8).contiguous().to(D), norm[lo:hi].contiguous().to(D), thr.to(D))
    assert np.allclose(c.cpu().numpy(), np.array(want["sums"]), rtol=1e-6, atol=1e-9), (c.cpu().numpy(), want["sums"])
    got = summarize_metrics(c)
    for k in ("PCK@0.05", "PCK@0.2", "mPCK", "NME", "AUC", "EPE"):
        assert abs(got[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (k, got[k], want[k])
    assert got["samples"] == B


# ------------------------------------------------------------- tensor-core (tcgen05) GEMM
def test_split_f16_reconstructs_fp32():
    x = rnd(300, 588, seed=1, scale=2.0)
    x[0, :4] = torch.tensor([1e-6, -3e-5, 1000.0, 0.0])
    so = ops.split_f16(x.to(dev()))
    assert so.Kp == 640 and tuple(so.data.shape) == (300, 1280)
    d = so.data.float().cpu()
    rec = d[:, :588] + d[:, 640:640 + 588]
    assert (d[:, 588:640] == 0).all() and (d[:, 640 + 588:] == 0).all()
    err = (rec - x).abs()
    bound = torch.maximum(x.abs() * 2.0 ** -21, torch.full_like(x, 2.0 ** -24))
    assert (err <= bound).all(), f"split error {err.max().item():.3e}"
    # strided 3-D view (cls row dropped) read in place
    t = rnd(3, 17, 64, seed=2).to(dev())
    so = ops.split_f16(t[:, 1:, :])
    rec = (so.data[:, :64].float() + so.data[:, 64:].float()).cpu()
    close(rec, t[:, 1:, :].reshape(48, 64), tol=1e-6, what="split seg")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (128, 256, 192), (1300, 768, 768),
                                   (200, 96, 100), (5184, 256, 768), (650, 3072, 768), (650, 768, 3072), (77, 33, 516)])
@pytest.mark.parametrize("tile_n", [128, 256, 512])
def test_gemm_tc_fp32_grade(M, N, K, tile_n, request):
    from edgecape_b200 import _lib
    _lib.load().ec_tc_set_tile_n(tile_n)
    request.addfinalizer(lambda: _lib.load().ec_tc_set_tile_n(0))
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.02), rnd(N, seed=3)
    want = (x.double() @ w.double().T + b.double())
    D = dev()
    got = ops.gemm_tc(ops.split_f16(x.to(D)), ops.split_f16(w.to(D), 2.0 ** 10), bias=b.to(D))
    # fp32-grade: the TMEM accumulator rounds toward zero once per UMMA (3*K/16 roundings), so the bound
    # grows with K; 2e-5 of the output scale covers K = 3072 (fc2) -- 50x inside the 1e-3 parity bar
    err = close(got, want.float(), tol=2e-5, what=f"tc {M}x{N}x{K}")
    ref = ops.gemm(x.to(D), w.to(D), bias=b.to(D))
    e32 = (ref.double().cpu() - want).abs().max().item() / want.abs().max().item()
    print(f"tc {M}x{N}x{K}: tcgen05 split-fp16 err {err:.2e}, fp32 FFMA err {e32:.2e}")


def test_gemm_tc_epilogues_views_and_split_out():
    M, N, K = 700, 320, 256
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    g, r = rnd(N, seed=4), rnd(M, N, seed=5)
    D = dev()
    a2, b2 = ops.split_f16(x.to(D)), ops.split_f16(w.to(D), 64.0)
    y = F.linear(x, w, b)
    for act, fn in ((ops.ACT_NONE, lambda t: t), (ops.ACT_RELU, F.relu), (ops.ACT_GELU, F.gelu), (ops.ACT_TANH, torch.tanh)):
        got = ops.gemm_tc(a2, b2, bias=b.to(D), act=act, colscale=g.to(D), residual=r.to(D))
        close(got, r + g * fn(y), tol=1e-5, what=f"tc act{act}")
    got = ops.gemm_tc(a2, b2, bias=b.to(D), act=ops.ACT_TANH, residual=r.to(D), res_mode=ops.RES_GATE)
    close(got, (torch.tanh(y) + 1) * r, tol=1e-5, what="tc gate")
    # in-place residual, column-slice output view, batch-strided 3-D output view
    acc = r.to(D).clone()
    ops.gemm_tc(a2, b2, bias=b.to(D), residual=acc, out=acc)
    close(acc, r + y, tol=1e-5, what="tc in-place")
    wide = torch.zeros(M, 2 * N, device=D)
    ops.gemm_tc(a2, b2, bias=b.to(D), out=wide[:, N:])
    close(wide[:, N:], y, tol=1e-5, what="tc ld view")
    assert (wide[:, :N] == 0).all()
    tok = torch.zeros(7, 101, N, device=D)
    ops.gemm_tc(a2, b2, bias=b.to(D), out=tok[:, 1:, :])
    close(tok[:, 1:, :].reshape(M, N), y, tol=1e-5, what="tc seg view")
    assert (tok[:, 0, :] == 0).all()
    # fused split of the result
    got, so = ops.gemm_tc(a2, b2, bias=b.to(D), act=ops.ACT_GELU, split_out=True)
    rec = so.data[:, :N].float() + so.data[:, so.Kp:so.Kp + N].float()
    close(rec, got, tol=1e-6, what="split_out")


def test_fused_split_outputs_of_layernorm_attention_and_gemm():
    D = dev()
    M, C = 650, 768
    x = rnd(M, C, seed=1, scale=2.0)
    w, b = 1 + 0.1 * rnd(C, seed=3), 0.1 * rnd(C, seed=4)
    want = F.layer_norm(x, (C,), w, b, 1e-6)
    y, so = ops.layernorm(x.to(D), w.to(D), b.to(D), 1e-6, split="also")
    so2 = ops.layernorm(x.to(D), w.to(D), b.to(D), 1e-6, split="only")
    close(y, want, what="ln fp32")
    for s_ in (so, so2):
        assert (s_.rows, s_.K, s_.Kp) == (M, C, C)
        close(s_.data[:, :C].float() + s_.data[:, C:].float(), want, tol=5e-6, what="ln split")
    B, H, L, Dh = 2, 12, 325, 64
    q, k, v = rnd(B, L, H * Dh, seed=5), rnd(B, L, H * Dh, seed=6), rnd(B, L, H * Dh, seed=7)
    o = ops.attention(q.to(D), k.to(D), v.to(D), H)
    sp = ops.attention(q.to(D), k.to(D), v.to(D), H, split="only")
    close(sp.data[:, :H * Dh].float() + sp.data[:, H * Dh:].float(), o.reshape(B * L, H * Dh), tol=2e-6, what="attn split")
    # split-only GEMM output feeding the next GEMM (fc1 -> GELU -> fc2 chain)
    w1, b1, w2 = rnd(512, C, seed=8, scale=0.03), rnd(512, seed=9), rnd(256, 512, seed=10, scale=0.04)
    _, h2 = ops.gemm_tc(so, ops.split_f16(w1.to(D), 256.0), bias=b1.to(D), act=ops.ACT_GELU, split_out=True,
                        fp32_out=False)
    got = ops.gemm_tc(h2, ops.split_f16(w2.to(D), 256.0))
    close(got, F.linear(F.gelu(F.linear(want, w1, b1)), w2), tol=2e-5, what="chained tc")


@pytest.mark.parametrize("B,H,Lq,Lk", [(2, 12, 325, 325), (3, 8, 100, 324), (1, 8, 324, 100), (2, 2, 128, 64),
                                       (1, 1, 5, 17), (2, 16, 130, 448), (1, 6, 257, 257)])
def test_attention_tc_matches_fp32_attention(B, H, Lq, Lk, monkeypatch):
    D_, E = 64, H * 64
    q, k, v = rnd(B, Lq, E, seed=1), rnd(B, Lk, E, seed=2), rnd(B, Lk, E, seed=3, scale=2.0)
    qh = q.double().view(B, Lq, H, D_).transpose(1, 2) * D_ ** -0.5
    kh = k.double().view(B, Lk, H, D_).transpose(1, 2)
    vh = v.double().view(B, Lk, H, D_).transpose(1, 2)
    want = ((qh @ kh.transpose(-1, -2)).softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    D = dev()
    monkeypatch.setattr(ops, "TENSOR_CORES", True)
    monkeypatch.setattr(ops, "ATTENTION_TC", True)
    names = []
    from edgecape_b200 import _lib
    orig = _lib.call
    monkeypatch.setattr(_lib, "call", lambda n, *a: (names.append(n), orig(n, *a))[1])
    got, sp = ops.attention(q.to(D), k.to(D), v.to(D), H, split="also")
    assert names[-1] == "ec_attention_tc"
    close(got, want, tol=5e-6, what=f"attention_tc {B},{H},{Lq},{Lk}")
    close(sp.data[:, :E].float() + sp.data[:, E:].float(), want.reshape(B * Lq, E), tol=5e-6, what="attention_tc split")
    monkeypatch.setattr(ops, "ATTENTION_TC", False)
    ref = ops.attention(q.to(D), k.to(D), v.to(D), H)
    assert names[-1] == "ec_attention"
    close(got, ref, tol=5e-6, what="tc vs simt attention")


@pytest.mark.parametrize("B,H,Lq,Lk,D_", [(2, 8, 424, 424, 32), (3, 8, 100, 100, 32), (2, 8, 100, 324, 64), (2, 4, 130, 200, 32)])
def test_attention_tc_with_mask_and_bias(B, H, Lq, Lk, D_, monkeypatch):
    """tcgen05 attention with key-padding mask and additive per-head bias, head dims 32 and 64 (encoder / biased decoder
    self-attention shapes), vs the fp64 statement and the fp32 SIMT kernel; one batch row is fully masked."""
    monkeypatch.setattr(ops, "TENSOR_CORES", True)
    monkeypatch.setattr(ops, "ATTENTION_TC", True)
    E = H * D_
    q, k, v = rnd(B, Lq, E, seed=1), rnd(B, Lk, E, seed=2), rnd(B, Lk, E, seed=3)
    bias = rnd(B, H, Lq, Lk, seed=4)
    mask = torch.zeros(B, Lk, dtype=torch.bool)
    mask[0, Lk - Lk // 3:] = True
    mask[0, 3] = True
    D = dev()
    for use_bias in (False, True):
        s = (q.double().view(B, Lq, H, D_).transpose(1, 2) * D_ ** -0.5) @ k.double().view(B, Lk, H, D_).transpose(1, 2).transpose(-1, -2)
        if use_bias:
            s = s + bias.double()
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
        want = (s.softmax(-1) @ v.double().view(B, Lk, H, D_).transpose(1, 2)).transpose(1, 2).reshape(B, Lq, E).float()
        got = ops.attention(q.to(D), k.to(D), v.to(D), H, key_mask=mask.to(torch.uint8).to(D),
                            bias=bias.to(D) if use_bias else None)
        close(got, want, tol=5e-6, what=f"attention_tc mask bias={use_bias} D={D_}")
    # a fully masked batch row gives zeros (as the fp32 kernel), not NaN
    full = torch.ones(B, Lk, dtype=torch.uint8)
    full[0] = 0
    got = ops.attention(q.to(D), k.to(D), v.to(D), H, key_mask=full.to(D))
    assert torch.isfinite(got).all() and (got[1:] == 0).all()


@pytest.mark.parametrize("B,N,H", [(2, 325, 12), (3, 257, 6), (1, 64, 1), (2, 130, 4), (1, 448, 2), (4, 17, 2), (1, 730, 16), (2, 449, 3), (1, 768, 2), (2, 577, 2)])
def test_attention_tma_on_split_qkv(B, N, H, monkeypatch):
    """TMA-fed tcgen05 attention on the split-fp16 output of the QKV GEMM (MN-major V operand)."""
    C = H * 64
    qkv = rnd(B * N, 3 * C, seed=11)
    q, k, v = (qkv[:, i * C:(i + 1) * C].double().view(B, N, H, 64).transpose(1, 2) for i in range(3))
    want = (((q / 8.0) @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(B, N, C).float()
    D = dev()
    qkv2 = ops.split_f16(qkv.to(D))
    got, sp = ops.attention_packed_split(qkv2, B, N, H, split="also")
    close(got, want, tol=5e-6, what=f"attention_tma {B},{N},{H}")
    close(sp.data[:, :C].float() + sp.data[:, C:].float(), want.reshape(B * N, C), tol=5e-6, what="attention_tma split")


@pytest.mark.parametrize("B,N,H", [(2, 325, 12), (1, 64, 1), (2, 130, 4), (1, 730, 16), (2, 449, 3)])
def test_attention_writes_f16f8_rows(B, N, H):
    """ec_attention_split_fmt_next(EC_SPLIT_F16F8): the split output rows in the [hi16 | per 64 columns: hi8, lo8] format ec_gemm_f16f8
    consumes (the ViT proj GEMM), from the persistent kernel (<= 368 tokens) and the one-tile-per-CTA kernel: the hi16
    plane is bit-identical to the F16X2 output, hi8 = e4m3(hi16), lo8 = e4m3(lo 2^11); then the proj-shaped GEMM on it."""
    C = H * 64
    D = dev()
    qkv2 = ops.split_f16(rnd(B * N, 3 * C, seed=21).to(D))
    ref = ops.attention_packed_split(qkv2, B, N, H)
    got = ops.attention_packed_split(qkv2, B, N, H, out_fmt=ops.F16F8)
    assert got.fmt == ops.F16F8 and got.Kp == C
    hi16, lo16 = ref.data[:, :C], ref.data[:, C:].float()
    raw = got.data.view(torch.uint8).view(B * N, 4 * C)
    assert torch.equal(raw[:, :2 * C].contiguous().view(torch.float16), hi16)
    e = raw[:, 2 * C:].reshape(B * N, H, 2, 64)                     # per 64 columns (= per head): [hi8 x 64 | lo8 x 64]
    h8 = e[:, :, 0].reshape(B * N, C).contiguous().view(torch.float8_e4m3fn).float()
    l8 = e[:, :, 1].reshape(B * N, C).contiguous().view(torch.float8_e4m3fn).float()
    assert torch.equal(h8, hi16.float().to(torch.float8_e4m3fn).float())
    # (the kernel rounds the exact fp32 low part, lo16 is that part rounded to fp16 first: at most one e4m3 step apart)
    exp_l8 = (lo16 * 2048.0).to(torch.float8_e4m3fn).float()
    assert bool(((l8 - exp_l8).abs() <= 0.13 * exp_l8.abs() + 2.0 ** -9).all())
    assert float((l8 != exp_l8).float().mean()) < 0.05
    # ... and as the A operand of the F16F8 GEMM against the F16X2 pair
    w = rnd(C, C, seed=22, scale=C ** -0.5).to(D)
    y8 = ops.gemm_tc(got, ops.split_weight(w, ops.F16F8))
    y16 = ops.gemm_tc(ref, ops.split_weight(w, ops.F16X2))
    close(y8, y16, tol=3e-5, what="proj on F16F8 attention rows")


@pytest.mark.parametrize("B,H,Lq,Lk,Dh,masked,biased", [
    (2, 8, 100, 100, 32, True, True),      # decoder self-attention over keypoints: fixed key mask + hop bias
    (2, 8, 424, 424, 32, True, False),     # encoder self-attention over image + keypoint tokens
    (3, 8, 100, 324, 64, False, False),    # keypoints -> image cross-attention
    (3, 8, 324, 100, 64, False, False),    # image -> keypoints cross-attention (two-way layers)
    (1, 8, 17, 448, 32, True, True),
    (2, 4, 130, 200, 64, True, True),
    (1, 2, 5, 3, 32, False, True),
])
def test_attention_split_general(B, H, Lq, Lk, Dh, masked, biased):
    """ec_attention_tc_split with separate Q / K / V operands, head dim 32 or 64, key mask and bias."""
    C = H * Dh
    q, k, v = rnd(B * Lq, C, seed=21), rnd(B * Lk, C, seed=22), rnd(B * Lk, C, seed=23)
    mask = None
    if masked:
        mask = (torch.rand(B, Lk, generator=torch.Generator().manual_seed(3)) < 0.3)
        mask[:, 0] = False
        if B > 1 and Lk > 3:
            mask[1, :] = True                     # a fully masked row block -> zeros
    bias = 0.5 * rnd(B, H, Lq, Lk, seed=24) if biased else None
    qh = q.double().view(B, Lq, H, Dh).transpose(1, 2)
    kh = k.double().view(B, Lk, H, Dh).transpose(1, 2)
    vh = v.double().view(B, Lk, H, Dh).transpose(1, 2)
    s_ = qh @ kh.transpose(-1, -2) * Dh ** -0.5
    if bias is not None:
        s_ = s_ + bias.double()
    if mask is not None:
        s_ = s_.masked_fill(mask[:, None, None, :], float("-inf"))
    p_ = torch.nan_to_num(s_.softmax(-1), nan=0.0)
    want = (p_ @ vh).transpose(1, 2).reshape(B, Lq, C).float()
    D = dev()
    q2, k2, v2 = ops.split_f16(q.to(D)), ops.split_f16(k.to(D)), ops.split_f16(v.to(D))
    got, sp = ops.attention_split(q2, 0, Lq, k2, 0, v2, 0, Lk, B, H, Lq, Lk, Dh,
                                  key_mask=None if mask is None else mask.to(torch.uint8).to(D),
                                  bias=None if bias is None else bias.to(D), split="also")
    close(got, want, tol=5e-6, what=f"attention_split {B},{H},{Lq},{Lk},{Dh}")
    close(sp.data[:, :C].float() + sp.data[:, C:].float(), want.reshape(B * Lq, C), tol=5e-6, what="attention_split split")


def test_attention_tc_on_packed_qkv_views(monkeypatch):
    monkeypatch.setattr(ops, "TENSOR_CORES", True)
    monkeypatch.setattr(ops, "ATTENTION_TC", True)
    B, N, H, C = 4, 325, 12, 768
    qkv = rnd(B, N, 3 * C, seed=5).to(dev())
    got = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H)
    monkeypatch.setattr(ops, "ATTENTION_TC", False)
    ref = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H)
    close(got, ref, tol=5e-6, what="packed qkv")


def test_linear_dispatch_uses_tensor_cores_and_matches_simt(monkeypatch):
    from edgecape_b200 import _lib
    x, w, b = rnd(3, 324, 768, seed=1), rnd(256, 768, seed=2, scale=0.03), rnd(256, seed=3)
    D = dev()
    xd, wd, bd = x.to(D), torch.nn.Parameter(w.to(D), requires_grad=False), b.to(D)
    names = []
    orig = _lib.call
    monkeypatch.setattr(_lib, "call", lambda n, *a: (names.append(n), orig(n, *a))[1])
    monkeypatch.setattr(ops, "TENSOR_CORES", True)
    y_tc = ops.linear(xd, wd, bd)
    assert "ec_gemm_f16x3" in names and "ec_gemm" not in names
    monkeypatch.setattr(ops, "TENSOR_CORES", False)
    y_32 = ops.linear(xd, wd, bd)
    assert names[-1] == "ec_gemm"
    close(y_tc, y_32, tol=3e-6, what="tc vs simt")
    close(y_tc, F.linear(x, w, b), tol=3e-6, what="tc vs torch")


@pytest.mark.parametrize("B,K,H", [(16, 100, 4), (3, 17, 4), (2, 200, 4), (1, 5, 2), (4, 33, 1)])
def test_markov_powers_one_launch(B, K, H):
    """ec_markov_powers: planes 2..H of the hop tensor = torch.matrix_power(P, h) (skeleton.py:152-161), one launch."""
    g = torch.Generator().manual_seed(B * 1000 + K)
    P = torch.rand(B, K, K, generator=g)
    P = P / P.sum(-1, keepdim=True)
    hops = torch.zeros(H + 1, B, K, K)
    hops[0] = torch.eye(K)
    if H >= 1:
        hops[1] = P
    got = ops.markov_powers_(hops.to(dev())).cpu()
    for h in range(H + 1):
        want = torch.matrix_power(P.double(), h).float()
        assert (got[h] - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item()), h


@pytest.mark.parametrize("B,K,masked", [(3, 100, True), (2, 17, False), (16, 100, True), (1, 200, True)])
def test_hop_bias_fused_into_attention(B, K, masked):
    """ec_attention_hop_bias_next: the structural bias MLP (utils/bias_attn.py:188-191) evaluated inside the attention
    kernel equals the separate ec_hop_bias kernel + additive bias tensor, and the fp64 statement of the biased MHA."""
    H, Dh, n_hops, hidden = 8, 32, 5, 12
    g = torch.Generator().manual_seed(B * 100 + K)
    qkv = torch.randn(B * K, 3 * H * Dh, generator=g)
    P = torch.rand(B, K, K, generator=g)
    P = P / P.sum(-1, keepdim=True)
    hops = torch.stack([torch.matrix_power(P, h) for h in range(n_hops)])            # [n_hops, B, K, K]
    w0, b0 = torch.randn(hidden, n_hops, generator=g), torch.randn(hidden, generator=g) * 0.1
    w1, b1 = torch.randn(H, hidden, generator=g), torch.randn(H, generator=g) * 0.1
    mask = torch.zeros(B, K, dtype=torch.uint8)
    if masked:
        mask[:, K - K // 4:] = 1
    D_ = dev()
    qkv2 = ops.split_f16(qkv.to(D_))
    hp, km = hops.contiguous().to(D_), (mask.to(D_) if masked else None)
    ws = [t.to(D_) for t in (w0, b0, w1, b1)]
    fused = ops.attention_packed_split(qkv2, B, K, H, key_mask=km, hop=(hp, *ws), split="no").cpu()
    bias = ops.hop_bias(hp, *ws)
    sep = ops.attention_packed_split(qkv2, B, K, H, key_mask=km, bias=bias, split="no").cpu()
    # fp64 statement
    q, k, v = (qkv[:, i * H * Dh:(i + 1) * H * Dh].double().view(B, K, H, Dh).transpose(1, 2) for i in range(3))
    bias64 = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(hops.permute(1, 2, 3, 0).double(), w0.double(),
                                                                               b0.double())), w1.double(), b1.double())
    s = (q @ k.transpose(-1, -2)) * Dh ** -0.5 + bias64.permute(0, 3, 1, 2)
    s = s.masked_fill(mask.bool()[:, None, None, :], float("-inf"))
    want = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, K, H * Dh).float()
    scale = want.abs().max().item()
    assert (fused - want).abs().max().item() < 2e-5 * scale
    assert (fused - sep).abs().max().item() < 1e-5 * scale
