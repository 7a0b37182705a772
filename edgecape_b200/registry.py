"""mmpose / mmcv registry surface of the drop-in boundary.

The reference registers its classes into mmpose's `POSENETS` / `HEADS`, its own `TRANSFORMER`
registry (EdgeCape/models/utils/builder.py:5-14) and mmcv's `POSITIONAL_ENCODING`, and
`test.py:120` builds the model with `build_posenet(cfg.model)`.  When mmpose / mmcv are
importable this module registers into those very registries (force=True so it *replaces*
the reference's classes of the same name -- the drop-in); otherwise it provides a minimal
local registry with the same `register_module` / `build` behaviour, so `configs/test/*.py`
load unchanged either way.
"""
import inspect


class Registry:
    """Subset of mmcv.utils.Registry used by the reference."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls

        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)

    def __contains__(self, key):
        return key in self.module_dict

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    """mmcv.utils.build_from_cfg semantics: cfg['type'] names the class, the rest are kwargs."""
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise KeyError(f"config for the {registry.name} registry must be a dict with a `type` key: {cfg!r}")
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    if not inspect.isclass(cls):
        raise TypeError(f"type must be a str or class, got {type(cls)}")
    return cls(**args)


def _mm_registries():
    try:  # pragma: no cover - mmpose is not installed in the build image
        from mmpose.models.builder import HEADS, POSENETS
        from mmcv.cnn.bricks.transformer import POSITIONAL_ENCODING
        return POSENETS, HEADS, POSITIONAL_ENCODING, True
    except Exception:
        return Registry("posenet"), Registry("head"), Registry("position encoding"), False


POSENETS, HEADS, POSITIONAL_ENCODING, USING_MMPOSE = _mm_registries()
TRANSFORMER = Registry("Transformer")


def build_posenet(cfg):
    return POSENETS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_transformer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER, default_args)


def build_positional_encoding(cfg, default_args=None):
    return build_from_cfg(cfg, POSITIONAL_ENCODING, default_args)
