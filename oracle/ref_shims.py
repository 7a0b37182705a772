"""TEST INFRASTRUCTURE ONLY -- never imported by the product (edgecape_b200/).

Shim layer that lets the *unmodified* reference sources under /root/reference be
imported in the authoring container, where mmcv / mmpose / fairseq / fvcore are
not installed and torch.hub has no network.  It is used only by
``oracle/gen_golden.py`` (to freeze golden vectors) and by the optional
``tests/test_oracle_vs_reference.py`` cross-check; both skip when
/root/reference is absent (e.g. on the GPU box).

What is stubbed:
  mmcv.cnn.{Conv2d,Linear,xavier_init,...}, mmcv.cnn.bricks.{registry,transformer},
  mmcv.runner.BaseModule, mmcv.utils.{Registry,build_from_cfg}, mmcv.image,
  mmcv.visualization.image, mmpose.models(.builder), mmpose.models.detectors.base,
  mmpose.models.utils.ops.resize, mmpose.core.evaluation.keypoint_pck_accuracy,
  mmpose.core.post_processing.transform_preds, fairseq.utils.softmax,
  fairseq.modules.{fairseq_dropout,quant_noise}, fvcore.nn.weight_init.
None of these stubs contain reference code; they restate the public behaviour of
the (absent) third-party functions the reference calls.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("EDGECAPE_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "EdgeCape", "models"))


class _Registry:
    def __init__(self, name, **_):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self.module_dict[name or module.__name__] = module
            return module

        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls

        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def __contains__(self, key):
        return key in self.module_dict

    def build(self, cfg, default_args=None):
        return _build_from_cfg(cfg, self, default_args)


def _build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    return cls(**args)


def _xavier_init(module, gain=1, bias=0, distribution="normal"):
    if hasattr(module, "weight") and module.weight is not None:
        if distribution == "uniform":
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


def _keypoint_pck_accuracy(pred, gt, mask, thr, normalize):
    """mmpose 0.29 `keypoint_pck_accuracy` restated: per-keypoint-type accuracy
    over the batch, then mean over keypoint types with >=1 valid sample."""
    N, K, _ = pred.shape
    dist = np.full((K, N), -1.0, dtype=np.float32)
    nrm = normalize.copy().astype(np.float32)
    nrm[np.where(nrm <= 0)] = 1e6
    valid = mask.copy()
    valid[np.where((nrm == 0).sum(1))[0], :] = False
    d = np.linalg.norm(((pred - gt) / nrm[:, None, :])[valid], axis=-1)
    dist_t = dist.T
    dist_t[valid] = d
    dist = dist_t.T
    acc = np.zeros(K, dtype=np.float32)
    for k in range(K):
        v = dist[k] != -1
        acc[k] = (dist[k][v] < thr).sum() / v.sum() if v.sum() > 0 else -1
    valid_acc = acc[acc >= 0]
    cnt = len(valid_acc)
    avg = valid_acc.mean() if cnt > 0 else 0
    return acc, avg, cnt


def _transform_preds(coords, center, scale, output_size, use_udp=False):
    scale = np.asarray(scale) * 200.0
    if use_udp:
        sx = scale[0] / (output_size[0] - 1.0)
        sy = scale[1] / (output_size[1] - 1.0)
    else:
        sx = scale[0] / output_size[0]
        sy = scale[1] / output_size[1]
    out = coords.copy()
    out[:, 0] = coords[:, 0] * sx + center[0] - scale[0] * 0.5
    out[:, 1] = coords[:, 1] * sy + center[1] - scale[1] * 0.5
    return out


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_INSTALLED = {}


def install():
    """Install stubs + namespace packages and import the reference hot-path
    modules from REFERENCE_ROOT unmodified. Returns a dict of the modules."""
    if _INSTALLED:
        return _INSTALLED
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    POSITIONAL_ENCODING = _Registry("position encoding")
    TRANSFORMER_LAYER = _Registry("transformerLayer")
    TRANSFORMER_LAYER_SEQUENCE = _Registry("transformer-layers sequence")
    HEADS = _Registry("heads")
    POSENETS = _Registry("posenets")

    def _nyi(*a, **k):
        raise NotImplementedError("stubbed mmcv function (not on the hot path)")

    _mod("mmcv", imwrite=_nyi, imread=_nyi, imshow=_nyi)
    _mod("mmcv.cnn", Conv2d=nn.Conv2d, Linear=nn.Linear, xavier_init=_xavier_init,
         build_activation_layer=_nyi, build_conv_layer=_nyi, build_norm_layer=_nyi)
    _mod("mmcv.cnn.bricks")
    _mod("mmcv.cnn.bricks.registry", TRANSFORMER_LAYER=TRANSFORMER_LAYER,
         TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE)
    _mod("mmcv.cnn.bricks.transformer", POSITIONAL_ENCODING=POSITIONAL_ENCODING,
         build_positional_encoding=lambda cfg, default_args=None: _build_from_cfg(
             cfg, POSITIONAL_ENCODING, default_args),
         BaseTransformerLayer=_BaseModule, TransformerLayerSequence=_BaseModule,
         build_transformer_layer_sequence=_nyi)
    _mod("mmcv.runner", BaseModule=_BaseModule)
    _mod("mmcv.runner.base_module", BaseModule=_BaseModule)
    _mod("mmcv.utils", Registry=_Registry, build_from_cfg=_build_from_cfg)
    _mod("mmcv.image", imwrite=_nyi)
    _mod("mmcv.visualization")
    _mod("mmcv.visualization.image", imshow=_nyi)

    builder = _mod("mmpose.models.builder", HEADS=HEADS, POSENETS=POSENETS,
                   build_head=lambda cfg: _build_from_cfg(cfg, HEADS),
                   build_posenet=lambda cfg: _build_from_cfg(cfg, POSENETS))
    _mod("mmpose")
    _mod("mmpose.models", HEADS=HEADS, POSENETS=POSENETS, builder=builder)
    _mod("mmpose.models.detectors")
    _mod("mmpose.models.detectors.base", BasePose=nn.Module)
    _mod("mmpose.models.utils")
    _mod("mmpose.models.utils.ops",
         resize=lambda input, size=None, scale_factor=None, mode="nearest",
         align_corners=None, warning=True: F.interpolate(
             input, size, scale_factor, mode, align_corners))
    _mod("mmpose.core")
    _mod("mmpose.core.evaluation", keypoint_pck_accuracy=_keypoint_pck_accuracy)
    _mod("mmpose.core.post_processing", transform_preds=_transform_preds)

    class FairseqDropout(nn.Module):
        def __init__(self, p, module_name=None):
            super().__init__()
            self.p = p

        def forward(self, x, inplace=False):
            return F.dropout(x, p=self.p, training=True, inplace=inplace) if (
                self.p > 0 and self.training) else x

    _mod("fairseq")
    _mod("fairseq.utils",
         softmax=lambda x, dim, onnx_trace=False: F.softmax(x, dim=dim, dtype=torch.float32))
    _mod("fairseq.modules")
    _mod("fairseq.modules.fairseq_dropout", FairseqDropout=FairseqDropout)
    _mod("fairseq.modules.quant_noise", quant_noise=lambda module, p, block_size: module)
    _mod("fvcore")
    _mod("fvcore.nn")
    _mod("fvcore.nn.weight_init", c2_xavier_fill=_nyi, c2_msra_fill=_nyi)

    # namespace packages so EdgeCape/__init__.py (datasets -> xtcocotools) never runs
    def _pkg(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        sys.modules[name] = m
        return m

    _pkg("EdgeCape", "EdgeCape")
    _pkg("EdgeCape.models", "EdgeCape/models")
    _pkg("EdgeCape.models.keypoint_heads", "EdgeCape/models/keypoint_heads")
    _pkg("EdgeCape.models.detectors", "EdgeCape/models/detectors")
    _pkg("EdgeCape.models.backbones", "EdgeCape/models/backbones")

    import importlib
    utils = importlib.import_module("EdgeCape.models.utils")
    encdec = importlib.import_module("EdgeCape.models.keypoint_heads.encoder_decoder")
    skeleton = importlib.import_module("EdgeCape.models.keypoint_heads.skeleton")
    head = importlib.import_module("EdgeCape.models.keypoint_heads.head")
    detector = importlib.import_module("EdgeCape.models.detectors.EdgeCape")
    posenc = importlib.import_module("EdgeCape.models.utils.positional_encoding")
    bias_attn = importlib.import_module("EdgeCape.models.utils.bias_attn")

    _INSTALLED.update(dict(utils=utils, encoder_decoder=encdec, skeleton=skeleton, head=head,
                           detector=detector, positional_encoding=posenc, bias_attn=bias_attn,
                           HEADS=HEADS, POSENETS=POSENETS,
                           POSITIONAL_ENCODING=POSITIONAL_ENCODING))
    return _INSTALLED


class AttrDict(dict):
    """dict with attribute access (stand-in for mmcv ConfigDict)."""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_attr(v) for v in obj)
    return obj


def build_reference_detector(model_cfg, backbone_module):
    """Build the reference `EdgeCape` detector with `torch.hub.load` patched to
    return `backbone_module` (an object exposing get_intermediate_layers)."""
    mods = install()
    cfg = to_attr(model_cfg)
    orig = torch.hub.load
    torch.hub.load = lambda repo, name, **kw: backbone_module
    try:
        args = dict(cfg)
        args.pop("type", None)
        det = mods["detector"].EdgeCape(**args)
    finally:
        torch.hub.load = orig
    det.eval()
    return det
