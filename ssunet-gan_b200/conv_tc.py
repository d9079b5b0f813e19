"""Host side of the tcgen05/TMEM/TMA implicit-GEMM convolution kernels (csrc/conv_tc.cu)."""
import torch

from . import _lib
from ._lib import W_RSCK_FLIP, W_RSKC, call


def eligible(cin, cout, k, stride):
    """Shapes the tensor-core forward kernel takes: 64-channel input granularity, 1x1 / 3x3 same-size, stride 1."""
    return cin % 64 == 0 and cin >= 64 and k in (1, 3) and stride == 1


def _flops(n, h, w, cin, cout, k):
    return 2.0 * n * h * w * cin * cout * k * k


def forward(x, weight, bias, y, stride, pad, act, slope, x1=None):
    from .ops import packed_weight
    n, c0, h, w = x.shape
    c1 = x1.shape[1] if x1 is not None else 0
    cout, cin, k, _ = weight.shape
    assert cin == c0 + c1 and 2 * pad == k - 1 and stride == 1
    wp = packed_weight(weight, W_RSKC, torch.bfloat16)
    call("ssg_conv2d_fwd_tc", x, c0, x1, c1, wp, bias, y, n, h, w, cout, k, pad, act, slope, flops=_flops(n, h, w, cin, cout, k))


def dgrad(dy, weight, dx, stride, pad):
    """dx = conv(dy, flip(W)^T): the forward kernel with roles of cin / cout swapped."""
    from .ops import packed_weight
    n, cout, h, w = dy.shape
    _, cin, k, _ = weight.shape
    assert stride == 1 and 2 * pad == k - 1
    wp = packed_weight(weight, W_RSCK_FLIP, torch.bfloat16)     # [tap'][cin][cout] == [taps][N][K]
    call("ssg_conv2d_fwd_tc", dy, cout, None, 0, wp, None, dx, n, h, w, cin, k, k - 1 - pad, 0, 0.0,
         flops=_flops(n, h, w, cin, cout, k))


def wgrad(x, dy, dw, stride, pad, x1=None):
    """dW (OIHW fp32) on tensor cores; returns False when the shape is not eligible (caller falls back to SIMT)."""
    n, c0, h, w = x.shape
    c1 = x1.shape[1] if x1 is not None else 0
    cout = dy.shape[1]
    k = dw.shape[-1]
    if stride != 1 or cout % 64 or c0 % 64 or c1 % 64 or k not in (1, 3) or 2 * pad != k - 1 or dy.dtype != torch.bfloat16:
        return False
    call("ssg_conv2d_wgrad_tc", x, c0, x1, c1, dy, dw, n, h, w, cout, k, pad, flops=_flops(n, h, w, c0 + c1, cout, k))
    return True
