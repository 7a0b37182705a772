"""CPU: the drop-in boundary -- C-ABI symbols, registries, config loading, state-dict keys, and the
"no CPU fallback" rule.  No kernel is launched here."""
import os

import pytest
import torch

import edgecape_b200 as E
from edgecape_b200 import _lib, registry
from edgecape_b200.config import default_model_cfg, state_dict_shapes, load_config
from edgecape_b200.synthetic import make_episode, make_state_dict

REF_CFG_DIR = "/root/reference/configs/test"
TINY_VIT = dict(embed_dim=64, depth=2, num_heads=4, patch_size=16, img_size=80)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"libedgecape_b200.so lacks {missing}"
    # and every symbol the Python side binds is declared in the public header
    assert set(_lib.SIGNATURES) <= set(declared)
    assert lib.ec_version() >= 100


def test_fused_gcn_shape_gate():
    """Host-side shape gate of the one-kernel GCN (no kernel is launched): slice width per (K, d, dff), 0 = the
    two-launch path takes the shape."""
    lib = _lib.load()
    assert lib.ec_gcn_fused_slice(100, 256, 384) == 192       # configs[1] decoder layers
    assert lib.ec_gcn_fused_slice(100, 256, 768) == 192       # skeleton layers at ViT-B (dff = C)
    assert lib.ec_gcn_fused_slice(100, 256, 1024) == 128      # ... at ViT-L
    assert lib.ec_gcn_fused_slice(64, 64, 64) == 64
    assert lib.ec_gcn_fused_slice(112, 256, 384) == 192
    assert lib.ec_gcn_fused_slice(128, 256, 384) == 64        # 16 KB tiles: only the narrow slice fits 227 KB
    assert lib.ec_gcn_fused_slice(200, 256, 384) == 0         # configs[4]: K > 128
    assert lib.ec_gcn_fused_slice(100, 192, 384) == 0         # d must be 64, 128 or 256
    assert lib.ec_gcn_fused_slice(100, 256, 100) == 0
    # the project-first kernel (the default): one CTA per item up to K = 128, a cluster of two CTAs up to K = 256
    assert lib.ec_gcn_fused2_slice(100, 256, 384) == 192
    assert lib.ec_gcn_fused2_slice(100, 256, 1024) == 128
    assert lib.ec_gcn_fused2_slice(128, 256, 384) == 128
    assert lib.ec_gcn_fused2_slice(200, 256, 384) == 128      # configs[4]: two CTAs of 100 rows, TMEM 2 x 128 + 208 columns
    assert lib.ec_gcn_fused2_slice(256, 256, 384) == 128
    assert lib.ec_gcn_fused2_slice(204, 256, 384) == 0        # K > 128 must be a multiple of 8
    assert lib.ec_gcn_fused2_slice(300, 256, 384) == 0
    assert lib.ec_gcn_fused2_slice(100, 192, 384) == 0


def test_registries_hold_the_reference_names():
    assert "EdgeCape" in registry.POSENETS
    assert "TwoStageHead" in registry.HEADS and "SkeletonPredictor" in registry.HEADS
    assert "TwoStageSupportRefineTransformer" in registry.TRANSFORMER
    assert "SinePositionalEncoding" in registry.POSITIONAL_ENCODING


def _tiny_cfg():
    cfg = default_model_cfg("dinov2_vits14")
    cfg["pretrained"] = TINY_VIT
    cfg["keypoint_head"]["in_channels"] = 64
    cfg["keypoint_head"]["skeleton_head"]["dim_feedforward"] = 64
    return cfg


def test_state_dict_keys_match_reference_layout():
    cfg = _tiny_cfg()
    model = E.build_model(dict(model=cfg))
    sd = model.state_dict()
    want = state_dict_shapes(cfg)
    assert set(sd) == set(want)
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    # one backbone bound to two names (detectors/EdgeCape.py:36)
    assert model.encoder_sample is model.encoder_query
    # a reference-style checkpoint loads strictly
    ck = make_state_dict(want, seed=3)
    res = model.load_state_dict(ck, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(model.state_dict()["keypoint_head_module.query_proj.weight"],
                       ck["keypoint_head_module.query_proj.weight"])


def test_forward_signature_matches_reference():
    import inspect
    sig = inspect.signature(E.EdgeCape.forward)
    assert list(sig.parameters)[1:9] == ["img_s", "img_q", "target_s", "target_weight_s", "target_q",
                                         "target_weight_q", "img_metas", "return_loss"]
    sig = inspect.signature(E.EdgeCape.forward_test)
    assert list(sig.parameters)[1:9] == ["img_s", "target_s", "target_weight_s", "img_q", "target_q",
                                         "target_weight_q", "img_metas", "vis_offset"]
    sig = inspect.signature(E.TwoStageHead.forward)
    assert list(sig.parameters)[1:6] == ["feature_q", "feature_s", "target_s", "mask_s", "skeleton_lst"]


@pytest.mark.skipif(not os.path.isdir(REF_CFG_DIR), reason="reference configs only exist in the authoring container")
@pytest.mark.parametrize("name", ["1shot_split1.py", "5shot_split1.py", "1shot_split3.py"])
def test_reference_test_configs_build_unchanged(name):
    cfg = load_config(os.path.join(REF_CFG_DIR, name))
    model = E.build_model(cfg)
    assert type(model).__name__ == "EdgeCape"
    assert model.keypoint_head_module.transformer.attn_bias is True
    assert model.keypoint_head_module.skeleton_head.learn_skeleton is True
    assert model.keypoint_head_module.skeleton_head.max_hop == 4


def test_no_cpu_fallback():
    """The product refuses to run on CPU tensors instead of silently falling back."""
    model = E.build_model(dict(model=_tiny_cfg()))
    data = make_episode(batch=1, image_size=64, num_kpts=5, skeleton="chain", seed=1)
    with pytest.raises(_lib.EdgeCapeLibraryError):
        model(return_loss=False, **data)
    with pytest.raises(NotImplementedError):
        model(return_loss=True, **data)
    from edgecape_b200 import ops
    with pytest.raises(_lib.EdgeCapeLibraryError):
        ops.gemm(torch.zeros(4, 4), torch.zeros(4, 4))


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.abspath(E.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_reference_style_checkpoint_with_fused_in_proj_loads():
    """BiasedMultiheadAttention accepts fused in_proj_{weight,bias} (utils/bias_attn.py:236-265)."""
    cfg = _tiny_cfg()
    sd = make_state_dict(state_dict_shapes(cfg), 5)
    fused = dict(sd)
    p = "keypoint_head_module.transformer.decoder.layers.0.self_attn."
    fused[p + "in_proj_weight"] = torch.cat([fused.pop(p + f"{n}_proj.weight") for n in "qkv"])
    fused[p + "in_proj_bias"] = torch.cat([fused.pop(p + f"{n}_proj.bias") for n in "qkv"])
    m1, m2 = E.build_model(dict(model=cfg)), E.build_model(dict(model=cfg))
    m1.load_state_dict(sd, strict=True)
    m2.load_state_dict(fused, strict=True)
    for (k1, v1), (k2, v2) in zip(sorted(m1.state_dict().items()), sorted(m2.state_dict().items())):
        assert k1 == k2 and torch.equal(v1, v2), k1
