"""EfficientNet encoder of the reference (scripts/efficientnet_pytorch/) on the hand-written sm_100a kernels."""
from .model import EfficientNet, MBConvBlock  # noqa: F401
from .utils import (  # noqa: F401
    BlockArgs,
    BlockDecoder,
    GlobalParams,
    efficientnet,
    get_model_params,
)

__version__ = "0.6.3"
