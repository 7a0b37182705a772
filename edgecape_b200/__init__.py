"""edgecape_b200 -- B200-native (sm_100a) implementation of EdgeCape's per-image inference hot path.

Importing the package registers the drop-in classes (`EdgeCape`, `TwoStageHead`,
`SkeletonPredictor`, `TwoStageSupportRefineTransformer`, `SinePositionalEncoding`) into the
mmpose / mmcv registries when those are importable, else into the local registries of
`edgecape_b200.registry`.  All arithmetic runs in `libedgecape_b200.so` (hand-written CUDA,
C ABI in include/edgecape_b200.h); there is no CPU or PyTorch-eager fallback.
"""
from . import registry  # noqa: F401
from .positional_encoding import SinePositionalEncoding  # noqa: F401
from .transformer import TwoStageSupportRefineTransformer  # noqa: F401
from .skeleton import SkeletonPredictor  # noqa: F401
from .head import TwoStageHead  # noqa: F401
from .detector import EdgeCape  # noqa: F401
from .vit import DinoVisionTransformerB200  # noqa: F401
from .config import build_model, default_model_cfg, load_config  # noqa: F401
from .registry import build_posenet  # noqa: F401

__version__ = "0.1.0"
