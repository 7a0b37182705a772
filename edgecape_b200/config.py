"""Config handling for the drop-in boundary.

The reference's configs (`configs/test/*.py`) are plain Python files that mmcv's
`Config.fromfile` executes; only their `model` dict reaches the hot path
(/root/reference/test.py:87,120).  `load_config` executes such a file with `runpy` (no mmcv
needed) and `build_model` hands `model` to the POSENETS registry, so the reference's files
load unchanged.  `default_model_cfg` is this repo's own statement of the stage-3 test model
(the values in /root/reference/configs/test/1shot_split1.py:32-72) for boxes where the
reference tree is absent.

One extension over the reference: `pretrained` may also be a dict
(`embed_dim, depth, num_heads, patch_size, img_size`) describing a custom ViT; the
reference only accepts hub names (`dinov2_vit{s,b,l}14`).
"""
import copy
import runpy

VIT_ARCHS = {
    # hub name: (embed_dim, depth, heads)   -- upstream dinov2/hub/backbones.py
    "dinov2_vits14": (384, 12, 6),
    "dinov2_vitb14": (768, 12, 12),
    "dinov2_vitl14": (1024, 24, 16),
}


def vit_config(pretrained):
    if isinstance(pretrained, dict):
        cfg = dict(patch_size=14, img_size=518, mlp_ratio=4, interpolate_offset=0.1)
        cfg.update(pretrained)
        return cfg
    if pretrained not in VIT_ARCHS:
        raise KeyError(f"unknown backbone {pretrained!r}; known: {sorted(VIT_ARCHS)} or a dict")
    dim, depth, heads = VIT_ARCHS[pretrained]
    return dict(embed_dim=dim, depth=depth, num_heads=heads, patch_size=14, img_size=518,
                mlp_ratio=4, interpolate_offset=0.1)


def default_model_cfg(pretrained="dinov2_vits14"):
    C = vit_config(pretrained)["embed_dim"]
    return dict(
        type="EdgeCape",
        pretrained=pretrained,
        encoder_config=dict(),
        keypoint_head=dict(
            type="TwoStageHead",
            in_channels=C,
            transformer=dict(
                type="TwoStageSupportRefineTransformer", d_model=256, nhead=8,
                num_encoder_layers=3, num_decoder_layers=3, dim_feedforward=384, dropout=0.1,
                similarity_proj_dim=256, dynamic_proj_dim=128, activation="relu",
                normalize_before=False, return_intermediate_dec=True, use_bias_attn_module=True,
                attn_bias=True, max_hops=4),
            share_kpt_branch=False, num_decoder_layer=3, with_heatmap_loss=False,
            heatmap_loss_weight=2.0, skeleton_loss_weight=1.0,
            positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True),
            skeleton_head=dict(type="SkeletonPredictor", learn_skeleton=True, dim_feedforward=C),
            learn_skeleton=True, masked_supervision=True, masking_ratio=0.5,
            model_freeze="skeleton"),
        train_cfg=dict(),
        test_cfg=dict(flip_test=False, post_process="default", shift_heatmap=True,
                      modulate_kernel=11),
    )


def load_config(path):
    """Execute an mmcv-style Python config file and return its namespace as a dict."""
    ns = runpy.run_path(path)
    return {k: v for k, v in ns.items() if not k.startswith("__")}


def build_model(cfg_or_path, **overrides):
    """`build_posenet(cfg.model)` equivalent (/root/reference/test.py:120)."""
    from .registry import build_posenet
    cfg = load_config(cfg_or_path) if isinstance(cfg_or_path, str) else cfg_or_path
    model = copy.deepcopy(cfg["model"] if "model" in cfg else cfg)
    model.update(overrides)
    return build_posenet(model)


# ------------------------------------------------------------------ state-dict key map
def _mha_keys(p, E, kdim=None, vdim=None):
    """torch.nn.MultiheadAttention parameter names."""
    if kdim is None:
        return {p + ".in_proj_weight": (3 * E, E), p + ".in_proj_bias": (3 * E,),
                p + ".out_proj.weight": (E, E), p + ".out_proj.bias": (E,)}
    return {p + ".q_proj_weight": (E, E), p + ".k_proj_weight": (E, kdim),
            p + ".v_proj_weight": (E, vdim), p + ".in_proj_bias": (3 * E,),
            p + ".out_proj.weight": (E, E), p + ".out_proj.bias": (E,)}


def _ln_keys(p, d):
    return {p + ".weight": (d,), p + ".bias": (d,)}


def _lin_keys(p, o, i):
    return {p + ".weight": (o, i), p + ".bias": (o,)}


def decoder_layer_shapes(p, d, nhead, dff, biased, max_hops, two_way, bias_mlp=True):
    """keypoint_heads/encoder_decoder.py:529-579."""
    s = {}
    if biased:
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s.update(_lin_keys(f"{p}.self_attn.{n}", d, d))
        if bias_mlp:    # torchvision.ops.MLP(max_hops+1, [max_hops+nhead, nhead]) -- utils/bias_attn.py:82
            s.update(_lin_keys(f"{p}.self_attn.markov_structural_mlp.0", max_hops + nhead, max_hops + 1))
            s.update(_lin_keys(f"{p}.self_attn.markov_structural_mlp.3", nhead, max_hops + nhead))
    else:
        s.update(_mha_keys(p + ".self_attn", d))
    s.update(_mha_keys(p + ".multihead_attn", 2 * d, 2 * d, d))
    s.update(_lin_keys(p + ".choker", d, 2 * d))
    s.update({p + ".ffn1.conv.weight": (2 * dff, d, 1), p + ".ffn1.conv.bias": (2 * dff,)})
    s.update(_lin_keys(p + ".ffn2", d, dff))
    for n in ("norm1", "norm2", "norm3"):
        s.update(_ln_keys(f"{p}.{n}", d))
    if two_way:
        s.update(_mha_keys(p + ".cross_attn_image_to_token", 2 * d, 2 * d, d))
        s.update(_lin_keys(p + ".cross_attn_image_to_token_choker", d, 2 * d))
        s.update(_ln_keys(p + ".norm4", d))
    return s


def head_shapes(hcfg, p="keypoint_head_module"):
    """State-dict keys/shapes of TwoStageHead (keypoint_heads/head.py:68-141)."""
    t = hcfg["transformer"]
    d, nhead = t.get("d_model", 256), t.get("nhead", 8)
    dff = t.get("dim_feedforward", 2048)
    C = hcfg["in_channels"]
    s = {f"{p}.transformer.mask_token": (1, d)}
    for i in range(t.get("num_encoder_layers", 3)):
        q = f"{p}.transformer.encoder.layers.{i}"
        s.update(_mha_keys(q + ".self_attn", d))
        s.update(_lin_keys(q + ".linear1", dff, d))
        s.update(_lin_keys(q + ".linear2", d, dff))
        s.update(_ln_keys(q + ".norm1", d))
        s.update(_ln_keys(q + ".norm2", d))
    biased = t.get("attn_bias", False) or t.get("use_bias_attn_module", False)
    for i in range(t.get("num_decoder_layers", 3)):
        s.update(decoder_layer_shapes(f"{p}.transformer.decoder.layers.{i}", d, nhead, dff, biased,
                                      t.get("max_hops", 5), False, t.get("attn_bias", False)))
    s.update(_ln_keys(f"{p}.transformer.decoder.norm", d))
    s.update(_lin_keys(f"{p}.transformer.decoder.ref_point_head.layers.0", d, d))
    s.update(_lin_keys(f"{p}.transformer.decoder.ref_point_head.layers.1", d, d))
    pg = f"{p}.transformer.proposal_generator"
    pd, dd = t.get("similarity_proj_dim", 256), t.get("dynamic_proj_dim", 128)
    s.update(_lin_keys(pg + ".support_proj", pd, d))
    s.update(_lin_keys(pg + ".query_proj", pd, d))
    s.update(_lin_keys(pg + ".dynamic_proj.0", dd, d))
    s.update(_lin_keys(pg + ".dynamic_proj.2", d, dd))
    s.update({f"{p}.input_proj.weight": (d, C, 1, 1), f"{p}.input_proj.bias": (d,)})
    s.update(_lin_keys(f"{p}.query_proj", d, C))
    for i in range(hcfg.get("num_decoder_layer", 3)):
        for j in (0, 2, 4):
            s.update(_lin_keys(f"{p}.kpt_branch.{i}.mlp.{j}", d, d))
        s.update(_lin_keys(f"{p}.kpt_branch.{i}.mlp.6", 2, d))
    sk = dict(hcfg.get("skeleton_head") or {})
    sp = f"{p}.skeleton_head"
    sd_model, sn = sk.get("d_model", 256), sk.get("nhead", 8)
    sdff = sk.get("dim_feedforward", 384)
    for i in range(sk.get("num_layers", 3)):
        s.update(decoder_layer_shapes(f"{sp}.skeleton_predictor.{i}", sd_model, sn, sdff, False,
                                      sk.get("max_hops", 4), sk.get("two_way_attn", True)))
    s.update({sp + ".image_project.weight": (sd_model, sdff, 1, 1), sp + ".image_project.bias": (sd_model,)})
    s.update(_lin_keys(sp + ".k_proj", sd_model, sd_model))
    s.update(_lin_keys(sp + ".q_proj", sd_model, sd_model))
    s.update({sp + ".mh_linear.weight": (1, sn, 1, 1), sp + ".mh_linear.bias": (1,)})
    if sk.get("use_zero_conv", True):
        s.update({sp + ".zero_conv.weight": (1, 1, 1, 1), sp + ".zero_conv.bias": (1,)})
    return s


def vit_shapes(vcfg, p):
    """Upstream DinoVisionTransformer state-dict keys (dinov2/models/vision_transformer.py)."""
    C, P = vcfg["embed_dim"], vcfg["patch_size"]
    G = vcfg["img_size"] // P
    Hd = int(C * vcfg["mlp_ratio"])
    s = {p + "cls_token": (1, 1, C), p + "pos_embed": (1, 1 + G * G, C), p + "mask_token": (1, C),
         p + "patch_embed.proj.weight": (C, 3, P, P), p + "patch_embed.proj.bias": (C,),
         p + "norm.weight": (C,), p + "norm.bias": (C,)}
    for i in range(vcfg["depth"]):
        b = f"{p}blocks.{i}."
        s.update({b + "norm1.weight": (C,), b + "norm1.bias": (C,),
                  b + "attn.qkv.weight": (3 * C, C), b + "attn.qkv.bias": (3 * C,),
                  b + "attn.proj.weight": (C, C), b + "attn.proj.bias": (C,),
                  b + "ls1.gamma": (C,),
                  b + "norm2.weight": (C,), b + "norm2.bias": (C,),
                  b + "mlp.fc1.weight": (Hd, C), b + "mlp.fc1.bias": (Hd,),
                  b + "mlp.fc2.weight": (C, Hd), b + "mlp.fc2.bias": (C,),
                  b + "ls2.gamma": (C,)})
    return s


def state_dict_shapes(model_cfg):
    """Every key of the detector's state dict: `encoder_sample.*` and `encoder_query.*` (one
    backbone bound to two names, detectors/EdgeCape.py:36) plus `keypoint_head_module.*`."""
    vcfg = vit_config(model_cfg.get("pretrained", "dinov2_vits14"))
    s = {}
    s.update(vit_shapes(vcfg, "encoder_sample."))
    s.update(vit_shapes(vcfg, "encoder_query."))
    s.update(head_shapes(model_cfg["keypoint_head"]))
    return s
