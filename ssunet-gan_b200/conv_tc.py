"""Host side of the tcgen05/TMEM/TMA implicit-GEMM convolution kernels (csrc/conv_tc.cu)."""
import torch

from . import _lib
from ._lib import W_RSCK, W_RSKC, call


def eligible(cin, cout, k, stride):
    """Shapes the tensor-core forward kernel takes: 64-channel input granularity, 1x1 / 3x3, stride 1 or 2."""
    return cin % 64 == 0 and cin >= 64 and k in (1, 3) and stride in (1, 2)


def dgrad_eligible(cin, cout, k, stride):
    return cout % 64 == 0 and cout >= 64 and k in (1, 3) and (stride == 1 or (stride == 2 and k == 3))


def _out_hw(h, w, k, stride, pad):
    return (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1


def _flops(n, oh, ow, cin, cout, k):
    return 2.0 * n * oh * ow * cin * cout * k * k


def forward(x, weight, bias, y, stride, pad, act, slope, x1=None):
    from .ops import packed_weight
    n, c0, h, w = x.shape
    c1 = x1.shape[1] if x1 is not None else 0
    cout, cin, k, _ = weight.shape
    assert cin == c0 + c1
    oh, ow = _out_hw(h, w, k, stride, pad)
    wp = packed_weight(weight, W_RSKC, torch.bfloat16)
    call("ssg_conv2d_fwd_tc", x, c0, x1, c1, wp, bias, y, n, h, w, cout, k, stride, pad, act, slope,
         flops=_flops(n, oh, ow, cin, cout, k))


def dgrad(dy, weight, dx, stride, pad):
    """dx (n, cin, h, w) from dy (n, cout, oh, ow); weights packed [tap][cin][cout] (K-major in cout)."""
    from .ops import packed_weight
    n, cin, h, w = dx.shape
    cout, _, k, _ = weight.shape
    wp = packed_weight(weight, W_RSCK, torch.bfloat16)
    call("ssg_conv2d_dgrad_tc", dy, wp, dx, n, h, w, cin, cout, k, stride, pad,
         flops=_flops(n, dy.shape[2], dy.shape[3], cin, cout, k))


def wgrad(x, dy, dw, stride, pad, x1=None):
    """dW (OIHW fp32) on tensor cores; returns False when the shape is not eligible (caller falls back to SIMT)."""
    n, c0, h, w = x.shape
    c1 = x1.shape[1] if x1 is not None else 0
    cout = dy.shape[1]
    k = dw.shape[-1]
    if stride not in (1, 2) or cout % 64 or c0 % 64 or c1 % 64 or k not in (1, 3) or dy.dtype != torch.bfloat16:
        return False
    call("ssg_conv2d_wgrad_tc", x, c0, x1, c1, dy, dw, n, h, w, cout, k, stride, pad,
         flops=_flops(n, dy.shape[2], dy.shape[3], c0 + c1, cout, k))
    return True
