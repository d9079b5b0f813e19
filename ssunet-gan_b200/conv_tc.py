"""Host side of the tcgen05/TMEM/TMA implicit-GEMM convolution kernels (csrc/conv_tc.cu)."""
import torch

from . import _lib
from ._lib import call, dtype_code

_ENABLED = None


def available():
    global _ENABLED
    if _ENABLED is None:
        _ENABLED = hasattr(_lib.lib(), "ssg_conv2d_fwd_tc")
    return _ENABLED


def eligible(cin, cout, k, stride):
    """Shapes the tensor-core kernels take: 64-channel granularity on both sides, 1x1 / 3x3, stride 1."""
    return available() and cin % 64 == 0 and cout % 64 == 0 and k in (1, 3) and stride == 1


def forward(x, weight, bias, y, stride, pad, act, slope):
    raise _lib.SsgError("tcgen05 convolution is not built into this library")


def dgrad(dy, weight, dx, stride, pad):
    raise _lib.SsgError("tcgen05 convolution is not built into this library")


def wgrad(x, dy, dw, stride, pad):
    return False
