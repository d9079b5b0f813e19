"""Model zoo entry points of the hot path (reference: archs.py).  `archs.__dict__[name]` is how the
reference instantiates networks (models_seg_gan.py:212-214, train.py:252-254); UNet_R_SS_v2 is config_v1's arch, the
other seven names of `archs.__all__` (archs.py:8) are SURVEY §8f row 4 and run on the same kernels.

Same constructor signatures, sub-module names, parameter registration order (=> identical
state_dict keys and identical default initialisation under the same torch seed) as the reference;
forward runs entirely on the hand-written kernels, activations in NHWC bf16 (or fp32)."""
import torch
from torch import nn
from torch.nn import init

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .nn_layers import BatchNorm2d, Conv2d, ReLU
from .normalization import SPADE

from .efficientnet_pytorch import EfficientNet
from .xresidualblock import xResidualBlock  # noqa: F401

__all__ = ["UNet", "NestedUNet", "SSUNet", "UNet_ori", "UNet_B_SS", "AttUNet", "UNet_R_SS", "UNet_R_SS_v2"]   # archs.py:8


def _fold_eval_bn(conv, bn):
    """Eval-mode BatchNorm folded into the convolution in front of it (inference only: archs.py:229-231 with running statistics):
    bn(conv(x, W)) = conv(x, W * s) + (beta - mean * s), s = gamma / sqrt(var + eps).  Cached on the conv module until a weight,
    a BN parameter or a running statistic changes (in-place updates by replayed graphs / fused optimisers bump ops._WEIGHT_EPOCH)."""
    key = (conv.weight._version, conv.weight.data_ptr(), bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, ops._WEIGHT_EPOCH)
    hit = getattr(conv, "_ssg_folded_bn", None)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    with torch.no_grad():
        s = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        w = (conv.weight.float() * s.view(-1, 1, 1, 1)).contiguous()
        b = (bn.bias.float() - bn.running_mean.float() * s).contiguous()
    conv._ssg_folded_bn = (key, w, b)
    return w, b


FOLD_EVAL_BN = True       # set False to keep the separate eval-mode BN pass (A/B switch for the inference leg)


class BasicBlock(nn.Module):
    """relu(bn2(conv2(relu(bn1(conv1 x)))) + shortcut(x))   (archs.py:205-241).
    BN-apply + ReLU and BN-apply + residual-add + ReLU are single fused passes."""
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = Conv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False))

    def init_weights(self):
        init.xavier_normal_(self.conv1.weight)
        init.xavier_normal_(self.conv2.weight)
        init.xavier_normal_(self.shortcut[0].weight)

    def forward(self, x):
        if isinstance(x, ops.CatPair):           # virtual concat [skip | upsampled]: both convolutions must be able to read it
            if not (len(self.shortcut) and ops._cat_conv_ok(x, self.conv1.weight, 1, 1) and self.conv1.stride == (1, 1)):
                x = x.materialise()
        else:
            x = ops.to_nhwc(x)
        # x feeds conv1 and the 1x1 shortcut: one gradient buffer for both (the second data-gradient kernel adds in its epilogue)
        sink = ops.grad_sink_for(x) if len(self.shortcut) else None
        if (FOLD_EVAL_BN and not self.training and not torch.is_grad_enabled() and self.bn1.running_mean is not None
                and self.bn1.weight is not None):
            # inference: bn1 + ReLU ride in conv1's epilogue (scaled weights, bias, activation) -- one activation pass less per block
            w1, b1 = _fold_eval_bn(self.conv1, self.bn1)
            out = ops.conv2d(x, w1, b1, self.conv1.stride[0], 1, ACT_RELU, 0.0)
        else:
            y1, s1 = self.conv1(x, want_stats=self.training, dx_sink=sink)  # BN statistics are reduced in the conv epilogue when possible
            out = self.bn1(y1, act=ACT_RELU, sums=s1)
        y2, s2 = self.conv2(out, want_stats=self.training)
        sc = self.shortcut[0](x, dx_sink=sink) if len(self.shortcut) else x
        return self.bn2(y2, residual=sc, act=ACT_RELU, sums=s2)


class AttentiveCNN(nn.Module):
    """EfficientNet feature extractor + 1x1 projection to 1024 channels (archs.py:409-466, `eff_flag` branch):
    bilinear resize to the encoder's native resolution -> `extract_features` -> `conv_a`.  Returns NCHW fp32."""

    _F_CHANNEL = {"efficientnet-b2": 1408, "efficientnet-b3": 1536, "efficientnet-b4": 1792, "efficientnet-b5": 2048}

    def __init__(self, model_info):
        super().__init__()
        self.f_channel = 1408
        eff_net_flag = model_info["eff_flag"]
        if eff_net_flag is not True:
            raise ops._lib.SsgError("AttentiveCNN: only the EfficientNet backbone (eff_flag=True) is on this package's path; "
                                    "the ResNet-101 branch (archs.py:436-443) needs torchvision's pretrained download")
        model_name = model_info["eff_model_name"]
        print("==> Building model.. : ", model_name)
        if model_info["phase_train"] is True:
            model = EfficientNet.from_pretrained(model_name, "../pretrained/normal/")
        else:
            model = EfficientNet.from_name(model_name)
        self.f_channel = self._F_CHANNEL.get(model_name, self.f_channel)
        self.eff_conv = model
        self.input_img_size = EfficientNet.get_image_size(model_name)
        self.eff_channel = 1024
        self.conv_a = Conv2d(self.f_channel, self.eff_channel, kernel_size=1, bias=False)
        self.eff_net_flag = eff_net_flag

    def forward(self, images):
        s = self.input_img_size
        resized = ops.resize_bilinear(ops.to_nhwc(images), s, s)
        return ops.to_nchw_f32(self.conv_a(self.eff_conv.extract_features(resized)))


class _Pool(nn.Module):
    """nn.MaxPool2d(2, 2, return_indices=True): returns (pooled, argmax code)."""

    def forward(self, x):
        return ops.max_pool2x2(x)


class _Unpool(nn.Module):
    def forward(self, x, code):
        return ops.max_unpool2x2(x, code)


class _Up(nn.Module):
    def forward(self, x):
        return ops.upsample_bilinear2x(x)


class UNet_R_SS_v2(nn.Module):
    """6-level residual U-Net with self-conditioned SPADE after every block (archs.py:559-671)."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        self.six_step = True
        nb_filter = [64, 128, 256, 384, 512, 768]
        spade_mid = num_classes
        self.pool = _Pool()
        self.unpool = _Unpool()
        self.up = _Up()
        context = "spadebatch3x3"
        ss_scale = 16
        f = nb_filter
        self.conv0_0 = BasicBlock(input_channels, f[0])
        self.SPADE0_0 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.conv1_0 = BasicBlock(f[0], f[1])
        self.SPADE1_0 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.conv2_0 = BasicBlock(f[1], f[2])
        self.SPADE2_0 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.conv3_0 = BasicBlock(f[2], f[3])
        self.SPADE3_0 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.conv4_0 = BasicBlock(f[3], f[4])
        self.SPADE4_0 = SPADE(context, f[4], spade_mid, f[4] / ss_scale)
        self.conv5_0 = BasicBlock(f[4], f[5])
        self.SPADE5_0 = SPADE(context, f[5], spade_mid, f[5] / ss_scale)
        self.conv_head5_0 = Conv2d(f[5], f[4], kernel_size=1, stride=1, bias=False)
        self.conv4_1 = BasicBlock(f[4] + f[4], f[4])
        self.SPADE4_1 = SPADE(context, f[4], spade_mid, f[4] / ss_scale)
        self.conv_head4_1 = Conv2d(f[4], f[3], kernel_size=1, stride=1, bias=False)
        self.conv3_1 = BasicBlock(f[3] + f[3], f[3])
        self.SPADE3_1 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.conv_head3_1 = Conv2d(f[3], f[2], kernel_size=1, stride=1, bias=False)
        self.conv2_1 = BasicBlock(f[2] + f[2], f[2])
        self.SPADE2_1 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.conv1_1 = BasicBlock(f[1] + f[2], f[1])
        self.SPADE1_1 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.conv0_1 = BasicBlock(f[0] + f[1], f[0])
        self.SPADE0_1 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.final = Conv2d(f[0], num_classes, kernel_size=1)
        self.init_weights()

    def init_weights(self):
        init.kaiming_uniform_(self.final.weight, mode="fan_in")
        self.final.bias.data.fill_(0)

    def forward(self, input):
        x = ops.to_nhwc(input, pad_channels=True)
        enc_0 = self.conv0_0(x)
        enc_0 = self.SPADE0_0(enc_0, enc_0)
        p0, _ = self.pool(enc_0)
        enc_1 = self.conv1_0(p0)
        enc_1 = self.SPADE1_0(enc_1, enc_1)
        p1, _ = self.pool(enc_1)
        enc_2 = self.conv2_0(p1)
        enc_2 = self.SPADE2_0(enc_2, enc_2)
        p2, i2 = self.pool(enc_2)
        enc_3 = self.conv3_0(p2)
        enc_3 = self.SPADE3_0(enc_3, enc_3)
        p3, i3 = self.pool(enc_3)
        enc_4 = self.conv4_0(p3)
        enc_4 = self.SPADE4_0(enc_4, enc_4)
        p4, i4 = self.pool(enc_4)
        enc_5 = self.conv5_0(p4)
        enc_5 = self.SPADE5_0(enc_5, enc_5)
        enc_5 = self.conv_head5_0(enc_5)
        dec_4 = self.conv4_1(ops.concat_channels(enc_4, self.unpool(enc_5, i4), virtual=True))
        dec_4 = self.SPADE4_1(dec_4, dec_4)
        dec_4 = self.conv_head4_1(dec_4)
        dec_3 = self.conv3_1(ops.concat_channels(enc_3, self.unpool(dec_4, i3), virtual=True))
        dec_3 = self.SPADE3_1(dec_3, dec_3)
        dec_3 = self.conv_head3_1(dec_3)
        dec_2 = self.conv2_1(ops.concat_channels(enc_2, self.unpool(dec_3, i2), virtual=True))
        dec_2 = self.SPADE2_1(dec_2, dec_2)
        dec_1 = self.conv1_1(ops.concat_channels(enc_1, self.up(dec_2), virtual=True))
        dec_1 = self.SPADE1_1(dec_1, dec_1)
        dec_0 = self.conv0_1(ops.concat_channels(enc_0, self.up(dec_1), virtual=True))
        dec_0 = self.SPADE0_1(dec_0, dec_0)
        nc = self.final.out_channels
        return ops.to_nchw_f32(self.final(dec_0, cout_store=ops.thin_pad(nc)), channels=nc)


# ----------------------------------------------------------------------------------------------
# the rest of archs.__all__: same kernels, other wiring
# ----------------------------------------------------------------------------------------------
def _pool(x):
    """nn.MaxPool2d(2, 2) without indices."""
    return ops.max_pool2x2(x)[0]


def _cat(*ts):
    """torch.cat(ts, 1) on NHWC storage."""
    out = ts[0]
    for t in ts[1:]:
        out = ops.concat_channels(out, t)
    return out


def _logits(conv, x):
    """The final 1x1 convolution to `num_classes` thin channels, returned as NCHW fp32 like the reference's output."""
    nc = conv.out_channels
    return ops.to_nchw_f32(conv(x, cout_store=ops.thin_pad(nc)), channels=nc)


def _conv_bn(conv, bn, x, training, residual=None, act=ACT_NONE):
    y, s = conv(x, want_stats=training)
    return bn(y, residual=residual, act=act, sums=s)


class _MaxPool2(nn.Module):
    def forward(self, x):
        return _pool(x)


class _UpNearest(nn.Module):
    """nn.Upsample(scale_factor=2): default mode is 'nearest'."""

    def forward(self, x):
        return ops.upsample_nearest2x(x)


class VGGBlock(nn.Module):
    """conv3x3(+bias) -> BN -> ReLU, twice (archs.py:94-112)."""

    def __init__(self, in_channels, middle_channels, out_channels):
        super().__init__()
        self.relu = ReLU()
        self.conv1 = Conv2d(in_channels, middle_channels, 3, padding=1)
        self.bn1 = BatchNorm2d(middle_channels)
        self.conv2 = Conv2d(middle_channels, out_channels, 3, padding=1)
        self.bn2 = BatchNorm2d(out_channels)

    def forward(self, x):
        out = _conv_bn(self.conv1, self.bn1, ops.to_nhwc(x), self.training, act=ACT_RELU)
        return _conv_bn(self.conv2, self.bn2, out, self.training, act=ACT_RELU)


class Bottleneck(nn.Module):
    """1x1 -> 3x3 -> 1x1 residual block with a conv+BN shortcut (archs.py:244-269; `expansion` is 1 there)."""
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = Conv2d(in_planes, planes, kernel_size=1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, self.expansion * planes, kernel_size=1, bias=False)
        self.bn3 = BatchNorm2d(self.expansion * planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False),
                                          BatchNorm2d(self.expansion * planes))

    def forward(self, x):
        x = ops.to_nhwc(x)
        out = _conv_bn(self.conv1, self.bn1, x, self.training, act=ACT_RELU)
        out = _conv_bn(self.conv2, self.bn2, out, self.training, act=ACT_RELU)
        sc = _conv_bn(self.shortcut[0], self.shortcut[1], x, self.training) if len(self.shortcut) else x
        return _conv_bn(self.conv3, self.bn3, out, self.training, residual=sc, act=ACT_RELU)


class conv_block(nn.Module):
    """archs.py:831-846: Sequential(conv3x3, BN, ReLU, conv3x3, BN, ReLU) under the attribute `conv`."""

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.conv = nn.Sequential(
            Conv2d(ch_in, ch_out, kernel_size=3, stride=1, padding=1, bias=True), BatchNorm2d(ch_out), ReLU(),
            Conv2d(ch_out, ch_out, kernel_size=3, stride=1, padding=1, bias=True), BatchNorm2d(ch_out), ReLU())

    def forward(self, x):
        c = self.conv
        out = _conv_bn(c[0], c[1], ops.to_nhwc(x), self.training, act=ACT_RELU)
        return _conv_bn(c[3], c[4], out, self.training, act=ACT_RELU)


class up_conv(nn.Module):
    """archs.py:848-861: Sequential(Upsample(x2 nearest), conv3x3, BN, ReLU) under the attribute `up`."""

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.up = nn.Sequential(_UpNearest(), Conv2d(ch_in, ch_out, kernel_size=3, stride=1, padding=1, bias=True),
                                BatchNorm2d(ch_out), ReLU())

    def forward(self, x):
        u = self.up
        return _conv_bn(u[1], u[2], u[0](ops.to_nhwc(x)), self.training, act=ACT_RELU)


class _Sigmoid(nn.Module):
    """Placeholder keeping `psi`'s Sequential indices; the sigmoid itself is fused into the gate kernel."""

    def forward(self, x):
        raise ops._lib.SsgError("Attention_block applies its sigmoid inside ops.pixel_gate")


class Attention_block(nn.Module):
    """x * sigmoid(BN(conv1x1(relu(BN(conv1x1 g) + BN(conv1x1 x)))))   (archs.py:115-142).
    BN + add + ReLU is one pass; the one-channel gate is applied (with its sigmoid) by one kernel."""

    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        self.W_g = nn.Sequential(Conv2d(F_g, F_int, kernel_size=1, stride=1, padding=0, bias=True), BatchNorm2d(F_int))
        self.W_x = nn.Sequential(Conv2d(F_l, F_int, kernel_size=1, stride=1, padding=0, bias=True), BatchNorm2d(F_int))
        self.psi = nn.Sequential(Conv2d(F_int, 1, kernel_size=1, stride=1, padding=0, bias=True), BatchNorm2d(1), _Sigmoid())
        self.relu = ReLU()

    def forward(self, g, x):
        g, x = ops.to_nhwc(g), ops.to_nhwc(x)
        g1 = _conv_bn(self.W_g[0], self.W_g[1], g, self.training)
        psi = _conv_bn(self.W_x[0], self.W_x[1], x, self.training, residual=g1, act=ACT_RELU)
        z = self.psi[1](self.psi[0](psi))
        return ops.pixel_gate(x, z)


class SubPixelConvolutionalBlock(nn.Module):
    """Parameter container only (archs.py:145-175): UNet_R_SS constructs `sp_up1_3` (archs.py:513) and never calls it, so
    its conv / PReLU parameters exist in the state_dict and receive no gradient."""

    def __init__(self, kernel_size=3, n_channels=64, scaling_factor=2):
        super().__init__()
        self.conv = Conv2d(in_channels=n_channels, out_channels=n_channels * (scaling_factor ** 2), kernel_size=kernel_size,
                           padding=kernel_size // 2)
        self.pixel_shuffle = nn.PixelShuffle(upscale_factor=scaling_factor)
        self.prelu = nn.PReLU()

    def forward(self, input):
        raise ops._lib.SsgError("SubPixelConvolutionalBlock.forward is not on any path of archs.__all__ (no pixel-shuffle kernel)")


class _PlainUNetBase(nn.Module):
    """The 5-level encoder/decoder wiring shared by UNet / SSUNet / UNet_B_SS / UNet_R_SS (archs.py:375-403,532-555,
    720-742,815-829): block -> [SPADE] -> pool ... ; decoder level = block(cat[skip, up(below)]) -> [SPADE]."""

    def _stage(self, name, x):
        x = getattr(self, "conv" + name)(x)
        sp = getattr(self, "SPADE" + name, None)
        return sp(x, x) if sp is not None else x

    def _forward5(self, input):
        x0_0 = self._stage("0_0", ops.to_nhwc(input, pad_channels=True))
        x1_0 = self._stage("1_0", _pool(x0_0))
        x2_0 = self._stage("2_0", _pool(x1_0))
        x3_0 = self._stage("3_0", _pool(x2_0))
        x4_0 = self._stage("4_0", _pool(x3_0))
        if getattr(self, "six_step", False):
            x5_0 = self._stage("5_0", _pool(x4_0))
            x4_1 = self._stage("4_1", _cat(x4_0, self.up(x5_0)))
            x3_1 = self._stage("3_1", _cat(x3_0, self.up(x4_1)))
        else:
            x3_1 = self._stage("3_1", _cat(x3_0, self.up(x4_0)))
        x2_2 = self._stage("2_2", _cat(x2_0, self.up(x3_1)))
        x1_3 = self._stage("1_3", _cat(x1_0, self.up(x2_2)))
        x0_4 = self._stage("0_4", _cat(x0_0, self.up(x1_3)))
        return x0_4, x1_3, x2_2, x3_1


class UNet(_PlainUNetBase):
    """archs.py:791-829."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        f = [64, 128, 256, 512, 1024]
        self.pool = _MaxPool2()
        self.up = _Up()
        self.conv0_0 = VGGBlock(input_channels, f[0], f[0])
        self.conv1_0 = VGGBlock(f[0], f[1], f[1])
        self.conv2_0 = VGGBlock(f[1], f[2], f[2])
        self.conv3_0 = VGGBlock(f[2], f[3], f[3])
        self.conv4_0 = VGGBlock(f[3], f[4], f[4])
        self.conv3_1 = VGGBlock(f[3] + f[4], f[3], f[3])
        self.conv2_2 = VGGBlock(f[2] + f[3], f[2], f[2])
        self.conv1_3 = VGGBlock(f[1] + f[2], f[1], f[1])
        self.conv0_4 = VGGBlock(f[0] + f[1], f[0], f[0])
        self.final = Conv2d(f[0], num_classes, kernel_size=1)

    def forward(self, input):
        return _logits(self.final, self._forward5(input)[0])


class ProgUNet(_PlainUNetBase):
    """archs.py:745-789 (not in `__all__`): the plain U-Net with one logits head per decoder level."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        f = [64, 128, 256, 512, 1024]
        self.pool = _MaxPool2()
        self.up = _Up()
        self.conv0_0 = VGGBlock(input_channels, f[0], f[0])
        self.conv1_0 = VGGBlock(f[0], f[1], f[1])
        self.conv2_0 = VGGBlock(f[1], f[2], f[2])
        self.conv3_0 = VGGBlock(f[2], f[3], f[3])
        self.conv4_0 = VGGBlock(f[3], f[4], f[4])
        self.conv3_1 = VGGBlock(f[3] + f[4], f[3], f[3])
        self.conv2_2 = VGGBlock(f[2] + f[3], f[2], f[2])
        self.conv1_3 = VGGBlock(f[1] + f[2], f[1], f[1])
        self.conv0_4 = VGGBlock(f[0] + f[1], f[0], f[0])
        self.final0 = Conv2d(f[0], num_classes, kernel_size=1)
        self.final1 = Conv2d(f[1], num_classes, kernel_size=1)
        self.final2 = Conv2d(f[2], num_classes, kernel_size=1)
        self.final3 = Conv2d(f[3], num_classes, kernel_size=1)

    def forward(self, input):
        x0_4, x1_3, x2_2, x3_1 = self._forward5(input)
        return [_logits(self.final0, x0_4), _logits(self.final1, x1_3), _logits(self.final2, x2_2), _logits(self.final3, x3_1)]


class SSUNet(_PlainUNetBase):
    """archs.py:673-742: VGG blocks of width [32..512], self-conditioned SPADE (hidden = C / 4) after every block."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        f = [32, 64, 128, 256, 512]
        spade_mid = num_classes
        self.pool = _MaxPool2()
        self.up = _Up()
        context = "spadebatch3x3"
        ss_scale = 4
        self.conv0_0 = VGGBlock(input_channels, f[0], f[0])
        self.SPADE0_0 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.conv1_0 = VGGBlock(f[0], f[1], f[1])
        self.SPADE1_0 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.conv2_0 = VGGBlock(f[1], f[2], f[2])
        self.SPADE2_0 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.conv3_0 = VGGBlock(f[2], f[3], f[3])
        self.SPADE3_0 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.conv4_0 = VGGBlock(f[3], f[4], f[4])
        self.SPADE4_0 = SPADE(context, f[4], spade_mid, f[4] / ss_scale)
        self.conv3_1 = VGGBlock(f[3] + f[4], f[3], f[3])
        self.SPADE3_1 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.conv2_2 = VGGBlock(f[2] + f[3], f[2], f[2])
        self.SPADE2_2 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.conv1_3 = VGGBlock(f[1] + f[2], f[1], f[1])
        self.SPADE1_3 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.conv0_4 = VGGBlock(f[0] + f[1], f[0], f[0])
        self.SPADE0_4 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.final = Conv2d(f[0], num_classes, kernel_size=1)

    def forward(self, input):
        return _logits(self.final, self._forward5(input)[0])


class UNet_B_SS(_PlainUNetBase):
    """archs.py:346-405: Bottleneck blocks of width [64..1024] + self-conditioned SPADE (all SPADEs registered first)."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        f = [64, 128, 256, 512, 1024]
        self.pool = _MaxPool2()
        self.up = _Up()
        context = "spadebatch3x3"
        ss_scale = 16
        spade_mid = num_classes
        self.SPADE0_0 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.SPADE1_0 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.SPADE2_0 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.SPADE3_0 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.SPADE4_0 = SPADE(context, f[4], spade_mid, f[4] / ss_scale)
        self.SPADE3_1 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.SPADE2_2 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.SPADE1_3 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.SPADE0_4 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.conv0_0 = Bottleneck(input_channels, f[0])
        self.conv1_0 = Bottleneck(f[0], f[1])
        self.conv2_0 = Bottleneck(f[1], f[2])
        self.conv3_0 = Bottleneck(f[2], f[3])
        self.conv4_0 = Bottleneck(f[3], f[4])
        self.conv3_1 = Bottleneck(f[3] + f[4], f[3])
        self.conv2_2 = Bottleneck(f[2] + f[3], f[2])
        self.conv1_3 = Bottleneck(f[1] + f[2], f[1])
        self.conv0_4 = Bottleneck(f[0] + f[1], f[0])
        self.final = Conv2d(f[0], num_classes, kernel_size=1)

    def forward(self, input):
        return _logits(self.final, self._forward5(input)[0])


class UNet_R_SS(_PlainUNetBase):
    """archs.py:469-555: the six-level residual SS-U-Net without unpooling (bilinear x2 at every decoder level)."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        self.six_step = True
        f = [64, 128, 256, 384, 512, 768]
        spade_mid = num_classes
        self.pool = _MaxPool2()
        self.up = _Up()
        context = "spadebatch3x3"
        ss_scale = 16
        self.conv0_0 = BasicBlock(input_channels, f[0])
        self.SPADE0_0 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.conv1_0 = BasicBlock(f[0], f[1])
        self.SPADE1_0 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.conv2_0 = BasicBlock(f[1], f[2])
        self.SPADE2_0 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.conv3_0 = BasicBlock(f[2], f[3])
        self.SPADE3_0 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.conv4_0 = BasicBlock(f[3], f[4])
        self.SPADE4_0 = SPADE(context, f[4], spade_mid, f[4] / ss_scale)
        self.conv5_0 = BasicBlock(f[4], f[5])
        self.SPADE5_0 = SPADE(context, f[5], spade_mid, f[5] / ss_scale)
        self.conv4_1 = BasicBlock(f[4] + f[5], f[4])
        self.SPADE4_1 = SPADE(context, f[4], spade_mid, f[4] / ss_scale)
        self.conv3_1 = BasicBlock(f[3] + f[4], f[3])
        self.SPADE3_1 = SPADE(context, f[3], spade_mid, f[3] / ss_scale)
        self.conv2_2 = BasicBlock(f[2] + f[3], f[2])
        self.SPADE2_2 = SPADE(context, f[2], spade_mid, f[2] / ss_scale)
        self.conv1_3 = BasicBlock(f[1] + f[2], f[1])
        self.SPADE1_3 = SPADE(context, f[1], spade_mid, f[1] / ss_scale)
        self.sp_up1_3 = SubPixelConvolutionalBlock(3, f[1], 2)
        self.conv0_4 = BasicBlock(f[0] + f[1], f[0])
        self.SPADE0_4 = SPADE(context, f[0], spade_mid, f[0] / ss_scale)
        self.final = Conv2d(f[0], num_classes, kernel_size=1)
        self.init_weights()

    def init_weights(self):
        init.kaiming_uniform_(self.final.weight, mode="fan_in")
        self.final.bias.data.fill_(0)

    def forward(self, input):
        return _logits(self.final, self._forward5(input)[0])


class NestedUNet(nn.Module):
    """U-Net++ (archs.py:863-933), optional deep supervision (a list of four logits tensors)."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        f = [64, 128, 256, 512, 1024]
        self.deep_supervision = deep_supervision
        self.pool = _MaxPool2()
        self.up = _Up()
        self.conv0_0 = VGGBlock(input_channels, f[0], f[0])
        self.conv1_0 = VGGBlock(f[0], f[1], f[1])
        self.conv2_0 = VGGBlock(f[1], f[2], f[2])
        self.conv3_0 = VGGBlock(f[2], f[3], f[3])
        self.conv4_0 = VGGBlock(f[3], f[4], f[4])
        self.conv0_1 = VGGBlock(f[0] + f[1], f[0], f[0])
        self.conv1_1 = VGGBlock(f[1] + f[2], f[1], f[1])
        self.conv2_1 = VGGBlock(f[2] + f[3], f[2], f[2])
        self.conv3_1 = VGGBlock(f[3] + f[4], f[3], f[3])
        self.conv0_2 = VGGBlock(f[0] * 2 + f[1], f[0], f[0])
        self.conv1_2 = VGGBlock(f[1] * 2 + f[2], f[1], f[1])
        self.conv2_2 = VGGBlock(f[2] * 2 + f[3], f[2], f[2])
        self.conv0_3 = VGGBlock(f[0] * 3 + f[1], f[0], f[0])
        self.conv1_3 = VGGBlock(f[1] * 3 + f[2], f[1], f[1])
        self.conv0_4 = VGGBlock(f[0] * 4 + f[1], f[0], f[0])
        if self.deep_supervision:
            self.final1 = Conv2d(f[0], num_classes, kernel_size=1)
            self.final2 = Conv2d(f[0], num_classes, kernel_size=1)
            self.final3 = Conv2d(f[0], num_classes, kernel_size=1)
            self.final4 = Conv2d(f[0], num_classes, kernel_size=1)
        else:
            self.final = Conv2d(f[0], num_classes, kernel_size=1)

    def forward(self, input):
        up = self.up
        x0_0 = self.conv0_0(ops.to_nhwc(input, pad_channels=True))
        x1_0 = self.conv1_0(_pool(x0_0))
        x0_1 = self.conv0_1(_cat(x0_0, up(x1_0)))
        x2_0 = self.conv2_0(_pool(x1_0))
        x1_1 = self.conv1_1(_cat(x1_0, up(x2_0)))
        x0_2 = self.conv0_2(_cat(x0_0, x0_1, up(x1_1)))
        x3_0 = self.conv3_0(_pool(x2_0))
        x2_1 = self.conv2_1(_cat(x2_0, up(x3_0)))
        x1_2 = self.conv1_2(_cat(x1_0, x1_1, up(x2_1)))
        x0_3 = self.conv0_3(_cat(x0_0, x0_1, x0_2, up(x1_2)))
        x4_0 = self.conv4_0(_pool(x3_0))
        x3_1 = self.conv3_1(_cat(x3_0, up(x4_0)))
        x2_2 = self.conv2_2(_cat(x2_0, x2_1, up(x3_1)))
        x1_3 = self.conv1_3(_cat(x1_0, x1_1, x1_2, up(x2_2)))
        x0_4 = self.conv0_4(_cat(x0_0, x0_1, x0_2, x0_3, up(x1_3)))
        if self.deep_supervision:
            return [_logits(self.final1, x0_1), _logits(self.final2, x0_2), _logits(self.final3, x0_3),
                    _logits(self.final4, x0_4)]
        return _logits(self.final, x0_4)


class _AttUNetBase(nn.Module):
    """Wiring shared by UNet_ori (archs.py:935-996) and AttUNet (archs.py:271-342): conv_block encoder, up_conv decoder,
    optional attention gate on each skip."""

    def _build(self, cin, cout, f, attention):
        self.Maxpool = _MaxPool2()
        self.Conv1 = conv_block(ch_in=cin, ch_out=f[0])
        self.Conv2 = conv_block(ch_in=f[0], ch_out=f[1])
        self.Conv3 = conv_block(ch_in=f[1], ch_out=f[2])
        self.Conv4 = conv_block(ch_in=f[2], ch_out=f[3])
        self.Conv5 = conv_block(ch_in=f[3], ch_out=f[4])
        for lvl in (5, 4, 3, 2):
            hi, lo = f[lvl - 1], f[lvl - 2]
            setattr(self, "Up%d" % lvl, up_conv(ch_in=hi, ch_out=lo))
            if attention:
                setattr(self, "Att%d" % lvl, Attention_block(F_g=lo, F_l=lo, F_int=lo // 2))
            setattr(self, "Up_conv%d" % lvl, conv_block(ch_in=hi, ch_out=lo))
        self.Conv_1x1 = Conv2d(f[0], cout, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        x1 = self.Conv1(ops.to_nhwc(x, pad_channels=True))
        x2 = self.Conv2(self.Maxpool(x1))
        x3 = self.Conv3(self.Maxpool(x2))
        x4 = self.Conv4(self.Maxpool(x3))
        d = self.Conv5(self.Maxpool(x4))
        for lvl, skip in ((5, x4), (4, x3), (3, x2), (2, x1)):
            d = getattr(self, "Up%d" % lvl)(d)
            att = getattr(self, "Att%d" % lvl, None)
            if att is not None:
                skip = att(d, skip)
            d = getattr(self, "Up_conv%d" % lvl)(_cat(skip, d))
        return _logits(self.Conv_1x1, d)


class UNet_ori(_AttUNetBase):
    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        self._build(input_channels, num_classes, [64, 128, 256, 512, 1024], attention=False)


class AttUNet(_AttUNetBase):
    def __init__(self, output_ch, img_ch=3, deep_supervision=False, **kwargs):
        super().__init__()
        self._build(img_ch, output_ch, [64, 128, 256, 512, 1024], attention=True)
