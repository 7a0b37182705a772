"""Parameter containers that reproduce the reference's state-dict key names.

The drop-in keeps `load_checkpoint(model, ckpt)` working: every module of this package owns
`nn.Parameter`s under exactly the names the reference's modules use
(`keypoint_head_module.transformer.encoder.layers.0.self_attn.in_proj_weight`, ...), built
from the {key: shape} maps in config.py.  Kernel-native repacked copies (fused QKV, packed GCN
weights, ...) are derived lazily and invalidated whenever parameters are loaded or moved.
"""
import math

import torch
import torch.nn as nn


class ParamTree(nn.Module):
    """Nested parameter container built from {dotted.key: shape}."""

    def __init__(self, shapes):
        super().__init__()
        groups = {}
        for key in sorted(shapes):
            head, _, rest = key.partition(".")
            if rest:
                groups.setdefault(head, {})[rest] = shapes[key]
            else:
                self.register_parameter(head, nn.Parameter(torch.zeros(tuple(shapes[key])), requires_grad=False))
        for name, sub in groups.items():
            self.add_module(name, ParamTree(sub))

    def __getitem__(self, key):
        """tree['layers.0.norm1.weight'] -> parameter tensor."""
        node = self
        for part in key.split("."):
            node = getattr(node, part)
        return node


def strip_prefix(shapes, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in shapes.items() if k.startswith(prefix)}


def xavier_uniform_all_(module):
    """The reference's init_weights: xavier-uniform on every >1-D `weight` (head.py:143-146)."""
    for name, p in module.named_parameters():
        if p.dim() > 1 and (name.endswith("weight") or "_weight" in name.rsplit(".", 1)[-1]):
            fan_out = p.shape[0] * (math.prod(p.shape[2:]) if p.dim() > 2 else 1)
            fan_in = p.shape[1] * (math.prod(p.shape[2:]) if p.dim() > 2 else 1)
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            with torch.no_grad():
                p.uniform_(-bound, bound)


class PackedMixin:
    """Lazy cache of kernel-native weight copies, dropped when parameters change."""

    def _init_packed(self):
        self._pk = None
        self.register_load_state_dict_post_hook(lambda mod, keys: mod.invalidate_packed())

    def invalidate_packed(self):
        self._pk = None
        for m in self.children():
            if isinstance(m, PackedMixin):
                m.invalidate_packed()

    def _apply(self, fn, *a, **k):
        self._pk = None
        return super()._apply(fn, *a, **k)

    def packed(self):
        if self._pk is None:
            with torch.no_grad():
                self._pk = self._pack()
        return self._pk

    def _pack(self):  # pragma: no cover - overridden
        return {}
