"""TEST INFRASTRUCTURE ONLY -- CPU restatement (PyTorch, fp32 or fp64) of the DINOv2
ViT that the reference obtains from ``torch.hub.load('facebookresearch/dinov2', ...)``
(/root/reference/EdgeCape/models/detectors/EdgeCape.py:35-36, called at :188-189 via
``get_intermediate_layers(img, n=1, reshape=True)[0]``).

DINOv2 is NOT vendored in /root/reference and not pinned (hub default branch).  This
file restates the published algorithm of upstream ``dinov2/models/vision_transformer.py``
(DinoVisionTransformer, `dinov2_vit{s,b,l}14`: patch 14, img_size 518, LayerNorm eps
1e-6, LayerScale, exact-erf GELU MLP x4, bicubic pos-embed interpolation with
``interpolate_offset=0.1`` and ``antialias=False``; no register tokens).
PARITY PIN: cross-checked in this container against ``transformers.Dinov2Model``
(an independent implementation of the same architecture) with shared random weights
-- see tests/test_oracle_dinov2.py.  There is no upstream golden vector: weights are
not available offline, so this boundary is "parity pinned against an independent
implementation, not against upstream outputs".

Resolution rule: inputs whose side is not a multiple of the patch
size follow *floor* semantics -- the stride-P conv ignores the trailing H % P rows /
columns (256 -> 18x18 patches), exactly what upstream computes once its PatchEmbed
shape assertion is lifted.

Only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() may import this module.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

ARCHS = {
    # name: (embed_dim, depth, heads)
    "dinov2_vits14": (384, 12, 6),
    "dinov2_vitb14": (768, 12, 12),
    "dinov2_vitl14": (1024, 24, 16),
}


def vit_config(pretrained):
    """Resolve a `pretrained` spec (hub name or explicit dict) into a config dict."""
    if isinstance(pretrained, dict):
        cfg = dict(patch_size=14, img_size=518, mlp_ratio=4, interpolate_offset=0.1)
        cfg.update(pretrained)
        return cfg
    dim, depth, heads = ARCHS[pretrained]
    return dict(embed_dim=dim, depth=depth, num_heads=heads, patch_size=14, img_size=518,
                mlp_ratio=4, interpolate_offset=0.1)


def vit_param_shapes(cfg):
    """Upstream state-dict keys and shapes of DinoVisionTransformer."""
    C, P = cfg["embed_dim"], cfg["patch_size"]
    G = cfg["img_size"] // P
    Hd = int(C * cfg["mlp_ratio"])
    shapes = {
        "cls_token": (1, 1, C),
        "pos_embed": (1, 1 + G * G, C),
        "mask_token": (1, C),
        "patch_embed.proj.weight": (C, 3, P, P),
        "patch_embed.proj.bias": (C,),
        "norm.weight": (C,),
        "norm.bias": (C,),
    }
    for i in range(cfg["depth"]):
        p = f"blocks.{i}."
        shapes.update({
            p + "norm1.weight": (C,), p + "norm1.bias": (C,),
            p + "attn.qkv.weight": (3 * C, C), p + "attn.qkv.bias": (3 * C,),
            p + "attn.proj.weight": (C, C), p + "attn.proj.bias": (C,),
            p + "ls1.gamma": (C,),
            p + "norm2.weight": (C,), p + "norm2.bias": (C,),
            p + "mlp.fc1.weight": (Hd, C), p + "mlp.fc1.bias": (Hd,),
            p + "mlp.fc2.weight": (C, Hd), p + "mlp.fc2.bias": (C,),
            p + "ls2.gamma": (C,),
        })
    return shapes


def interpolate_pos_embed(pos_embed, h0, w0, offset=0.1):
    """Upstream `interpolate_pos_encoding`: bicubic, scale_factor=((n+offset)/M), no antialias.
    pos_embed [1, 1+M*M, C] -> [1, 1+h0*w0, C]. Input independent (depends on h0,w0 only)."""
    N = pos_embed.shape[1] - 1
    M = int(math.sqrt(N))
    assert M * M == N
    if h0 * w0 == N and h0 == w0:
        return pos_embed
    dim = pos_embed.shape[-1]
    cls_pe = pos_embed[:, :1].float()
    patch_pe = pos_embed[:, 1:].float().reshape(1, M, M, dim).permute(0, 3, 1, 2)
    if offset:
        kw = dict(scale_factor=(float(h0 + offset) / M, float(w0 + offset) / M))
    else:
        kw = dict(size=(h0, w0))
    patch_pe = F.interpolate(patch_pe, mode="bicubic", antialias=False, **kw)
    assert patch_pe.shape[-2:] == (h0, w0), patch_pe.shape
    patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, h0 * w0, dim)
    return torch.cat((cls_pe, patch_pe), dim=1).to(pos_embed.dtype)


def vit_forward_tokens(sd, cfg, x, prefix=""):
    """Run the ViT; returns final-LayerNorm'ed patch tokens [B, h0*w0, C] (cls dropped)
    and (h0, w0).  `sd` maps upstream keys (with `prefix`) to tensors."""
    g = lambda k: sd[prefix + k]
    P, C, H = cfg["patch_size"], cfg["embed_dim"], cfg["num_heads"]
    B = x.shape[0]
    h0, w0 = x.shape[2] // P, x.shape[3] // P
    t = F.conv2d(x, g("patch_embed.proj.weight"), g("patch_embed.proj.bias"), stride=P)
    assert t.shape[-2:] == (h0, w0)
    t = t.flatten(2).transpose(1, 2)                                   # [B, S, C]
    t = torch.cat((g("cls_token").expand(B, -1, -1), t), dim=1)
    t = t + interpolate_pos_embed(g("pos_embed"), h0, w0, cfg.get("interpolate_offset", 0.1))
    d = C // H
    for i in range(cfg["depth"]):
        p = f"blocks.{i}."
        y = F.layer_norm(t, (C,), g(p + "norm1.weight"), g(p + "norm1.bias"), 1e-6)
        qkv = F.linear(y, g(p + "attn.qkv.weight"), g(p + "attn.qkv.bias"))
        qkv = qkv.reshape(B, -1, 3, H, d).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * d ** -0.5, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        y = (a @ v).transpose(1, 2).reshape(B, -1, C)
        y = F.linear(y, g(p + "attn.proj.weight"), g(p + "attn.proj.bias"))
        t = t + y * g(p + "ls1.gamma")
        y = F.layer_norm(t, (C,), g(p + "norm2.weight"), g(p + "norm2.bias"), 1e-6)
        y = F.gelu(F.linear(y, g(p + "mlp.fc1.weight"), g(p + "mlp.fc1.bias")))
        y = F.linear(y, g(p + "mlp.fc2.weight"), g(p + "mlp.fc2.bias"))
        t = t + y * g(p + "ls2.gamma")
    t = F.layer_norm(t, (C,), g("norm.weight"), g("norm.bias"), 1e-6)
    return t[:, 1:], (h0, w0)


class _LayerScale(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))


class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim)
        self.ls1 = _LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)
        self.ls2 = _LayerScale(dim)


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class DinoV2Oracle(nn.Module):
    """Parameter container with upstream state-dict key names plus the one method the
    reference calls, so it can stand in for the hub model inside the unmodified
    reference detector (oracle/ref_shims.py:build_reference_detector).  The arithmetic
    is `vit_forward_tokens` above."""

    def __init__(self, pretrained="dinov2_vits14"):
        super().__init__()
        cfg = self.cfg = vit_config(pretrained)
        C, P = cfg["embed_dim"], cfg["patch_size"]
        G = cfg["img_size"] // P
        self.embed_dim, self.patch_size = C, P
        self.cls_token = nn.Parameter(torch.zeros(1, 1, C))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + G * G, C))
        self.mask_token = nn.Parameter(torch.zeros(1, C))
        self.patch_embed = _PatchEmbed(C, P)
        self.blocks = nn.ModuleList(_Block(C, int(C * cfg["mlp_ratio"])) for _ in range(cfg["depth"]))
        self.norm = nn.LayerNorm(C, eps=1e-6)
        assert set(self.state_dict().keys()) == set(vit_param_shapes(cfg).keys())

    def get_intermediate_layers(self, x, n=1, reshape=False, return_class_token=False, norm=True):
        assert n == 1 and norm and not return_class_token
        tok, (h0, w0) = vit_forward_tokens(dict(self.state_dict(keep_vars=True)), self.cfg, x)
        if reshape:
            B = x.shape[0]
            tok = tok.reshape(B, h0, w0, -1).permute(0, 3, 1, 2).contiguous()
        return (tok,)
