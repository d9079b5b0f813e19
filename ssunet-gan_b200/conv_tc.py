"""Host side of the tcgen05/TMEM/TMA implicit-GEMM convolution kernels (csrc/conv_tc.cu)."""
import torch

from . import _lib
from ._lib import W_RSCK, W_RSKC, call


def eligible(cin_s, cout_s, k, stride):
    """Shapes the tensor-core kernels take: stored channel counts multiples of 8 (16-byte NHWC pixel pitch for TMA),
    1x1 / 3x3, stride 1 or 2 (1x1 only at stride 1)."""
    return cin_s % 8 == 0 and cout_s % 8 == 0 and (k == 3 or (k == 1 and stride == 1)) and stride in (1, 2)


def _out_hw(h, w, k, stride, pad):
    return (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1


def _flops(n, oh, ow, cin, cout, k):
    return 2.0 * n * oh * ow * cin * cout * k * k


def has_stats(k, stride, pad):
    """True when the forward kernel can emit BatchNorm sum / sum-of-squares as a by-product of its epilogue."""
    return bool(_lib.lib().ssg_conv2d_fwd_tc_has_stats(int(k), int(stride), int(pad)))


def forward(x, weight, bias, y, stride, pad, act, slope, x1=None, stats=None):
    """y = act(conv([x | x1], weight) + bias); x / x1 / y may be channel-padded (ops.thin_pad).
    stats: optional fp64 [2 * Cout_stored] tensor that receives per-channel sum(y), sum(y^2)."""
    from .ops import packed_weight
    n, c0, h, w = x.shape
    c1 = x1.shape[1] if x1 is not None else 0
    cout, cin, k, _ = weight.shape
    assert cin <= c0 + c1 and cout <= y.shape[1]
    oh, ow = _out_hw(h, w, k, stride, pad)
    wp = packed_weight(weight, W_RSKC, torch.bfloat16, cout_p=y.shape[1], cin_p=c0 + c1)
    call("ssg_conv2d_fwd_tc", x, c0, x1, c1, wp, bias, cout if bias is not None else 0, y, n, h, w, y.shape[1], k, stride, pad, act,
         slope, stats, flops=_flops(n, oh, ow, cin, cout, k))


def can_accumulate(k, stride, pad):
    """True when the data-gradient kernel can add into an existing dx (TMA reduce-add epilogue)."""
    return bool(_lib.lib().ssg_conv2d_dgrad_tc_can_acc(int(k), int(stride), int(pad)))


def can_mask(k, stride, pad):
    """True when the data-gradient kernel can fold the producer's activation backward into its epilogue."""
    return bool(_lib.lib().ssg_conv2d_dgrad_tc_mask_supported(int(k), int(stride), int(pad)))


def dgrad(dy, weight, dx, stride, pad, accumulate=False, producer_out=None, producer_act=0, producer_slope=0.0):
    """dx (n, cin_s, h, w) from dy (n, cout_s, oh, ow); weights packed [tap][cin_s][cout_s] (K-major in cout).
    accumulate: dx += ... (only for the geometries `can_accumulate` accepts).
    producer_out: the post-activation output of the layer that produced this convolution's input (== the saved input x): dx is
    multiplied by act'(producer_out) in the epilogue (only for the geometries `can_mask` accepts)."""
    from .ops import packed_weight
    n, cin_s, h, w = dx.shape
    cout, cin, k, _ = weight.shape
    cout_s = dy.shape[1]
    wp = packed_weight(weight, W_RSCK, torch.bfloat16, cout_p=cout_s, cin_p=cin_s)
    fl = _flops(n, dy.shape[2], dy.shape[3], cin, cout, k)
    if producer_out is not None:
        assert not accumulate and tuple(producer_out.shape) == tuple(dx.shape)
        call("ssg_conv2d_dgrad_tc_mask", dy, wp, dx, producer_out, int(producer_act), float(producer_slope), n, h, w, cin_s, cout_s, k,
             stride, pad, flops=fl)
        return
    call("ssg_conv2d_dgrad_tc_acc" if accumulate else "ssg_conv2d_dgrad_tc", dy, wp, dx, n, h, w, cin_s, cout_s, k, stride, pad, flops=fl)


def dgrad_split(dy, weight, dx0, dx1, stride, pad, accumulate=False):
    """Data gradient of a convolution over the virtual concatenation [x0 | x1]: gradient channels [0, c0) -> dx0, the rest -> dx1
    (written, or added when `accumulate`).  c0 % 64 == 0; only the geometries `can_accumulate` accepts."""
    from .ops import packed_weight
    n, c0, h, w = dx0.shape
    c1 = dx1.shape[1]
    cout, cin, k, _ = weight.shape
    assert cin == c0 + c1 and c0 % 64 == 0 and c1 % 8 == 0
    cout_s = dy.shape[1]
    wp = packed_weight(weight, W_RSCK, torch.bfloat16, cout_p=cout_s, cin_p=cin)
    call("ssg_conv2d_dgrad_tc_split", dy, wp, dx0, c0, dx1, c1, n, h, w, cout_s, k, stride, pad, int(bool(accumulate)),
         flops=_flops(n, dy.shape[2], dy.shape[3], cin, cout, k))


def wgrad(x, dy, dw, stride, pad, x1=None, accumulate=False):
    """dW (OIHW fp32, real channel extents) on tensor cores from (possibly channel-padded) x and dy.
    accumulate: dw += ... (dw is the parameter's slot of the flat gradient arena) instead of dw = ..."""
    n, c0, h, w = x.shape
    c1 = x1.shape[1] if x1 is not None else 0
    cout, cin, k, _ = dw.shape
    call("ssg_conv2d_wgrad_tc_acc" if accumulate else "ssg_conv2d_wgrad_tc", x, c0, x1, c1, dy, dy.shape[1], dw, cout, cin, n, h, w, k, stride, pad,
         flops=_flops(n, dy.shape[2], dy.shape[3], cin, cout, k))
    return True
