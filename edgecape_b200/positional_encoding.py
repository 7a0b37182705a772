"""@POSITIONAL_ENCODING SinePositionalEncoding -- same constructor and methods as
/root/reference/EdgeCape/models/utils/positional_encoding.py:11-122, evaluated by the
ec_sine_pe_coords kernel.  The grid encoding (`forward`, :57-94) is input independent for the
all-valid masks the head passes (head.py:172-173), so it is evaluated once per (h, w) and cached.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .registry import POSITIONAL_ENCODING


@POSITIONAL_ENCODING.register_module(force=True)
class SinePositionalEncoding(nn.Module):
    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi, eps=1e-6, offset=0.,
                 init_cfg=None):
        super().__init__()
        if normalize:
            assert isinstance(scale, (float, int)), \
                f"when normalize is set, scale should be provided and in float or int type, found {type(scale)}"
        self.num_feats = num_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = scale
        self.eps = eps
        self.offset = offset
        self._grid_cache = {}

    def grid_tokens(self, h, w, device):
        """Grid encoding for an all-valid h x w mask, token-major [h*w, 2*num_feats]."""
        key = (h, w, str(device))
        if key not in self._grid_cache:
            # coordinates exactly as the reference forms them in fp32 (cumsum of ones, :72-78);
            # the scale multiplication and everything after it happens in the kernel
            y = torch.arange(1, h + 1, dtype=torch.float32)
            x = torch.arange(1, w + 1, dtype=torch.float32)
            if self.normalize:
                y = (y + self.offset) / (y[-1:] + self.eps)
                x = (x + self.offset) / (x[-1:] + self.eps)
                scale = float(self.scale)
            else:
                scale = 1.0
            coord = torch.stack((x[None, :].expand(h, w), y[:, None].expand(h, w)), dim=-1).reshape(h * w, 2)
            coord = coord.contiguous().to(device)
            self._grid_cache[key] = ops.sine_pe_coords(coord, self.num_feats, float(self.temperature), scale)
        return self._grid_cache[key]

    def forward(self, mask):
        """mask [bs,h,w] (all zeros on the EdgeCape path) -> [bs, 2*num_feats, h, w] like the reference."""
        if bool(mask.any()):
            raise NotImplementedError("SinePositionalEncoding.forward: only all-valid masks are on the EdgeCape path")
        bs, h, w = mask.shape
        g = self.grid_tokens(h, w, mask.device if mask.is_cuda else torch.device("cuda"))
        return g.reshape(h, w, -1).permute(2, 0, 1)[None].expand(bs, -1, -1, -1)

    def forward_coordinates(self, coord):
        """coord [bs,kpt,2] in [0,1] -> [bs,kpt,2*num_feats] (:96-122)."""
        return ops.sine_pe_coords(coord.contiguous(), self.num_feats, float(self.temperature), float(self.scale))

    def __repr__(self):
        return (f"{self.__class__.__name__}(num_feats={self.num_feats}, temperature={self.temperature}, "
                f"normalize={self.normalize}, scale={self.scale}, eps={self.eps})")
