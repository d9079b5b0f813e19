"""Model zoo entry points of the hot path (reference: archs.py).  `archs.__dict__[name]` is how the
reference instantiates networks (models_seg_gan.py:212-214); UNet_R_SS_v2 is config_v1's arch.

Same constructor signatures, sub-module names, parameter registration order (=> identical
state_dict keys and identical default initialisation under the same torch seed) as the reference;
forward runs entirely on the hand-written kernels, activations in NHWC bf16 (or fp32)."""
import torch
from torch import nn
from torch.nn import init

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .nn_layers import BatchNorm2d, Conv2d
from .normalization import SPADE

from .efficientnet_pytorch import EfficientNet
from .xresidualblock import xResidualBlock  # noqa: F401

__all__ = ["UNet_R_SS_v2"]


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
        x = ops.to_nhwc(x)
        y1, s1 = self.conv1(x, want_stats=self.training)  # BN statistics are reduced in the conv epilogue when possible
        out = self.bn1(y1, act=ACT_RELU, sums=s1)
        y2, s2 = self.conv2(out, want_stats=self.training)
        sc = self.shortcut[0](x) if len(self.shortcut) else x
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
        dec_4 = self.conv4_1(ops.concat_channels(enc_4, self.unpool(enc_5, i4)))
        dec_4 = self.SPADE4_1(dec_4, dec_4)
        dec_4 = self.conv_head4_1(dec_4)
        dec_3 = self.conv3_1(ops.concat_channels(enc_3, self.unpool(dec_4, i3)))
        dec_3 = self.SPADE3_1(dec_3, dec_3)
        dec_3 = self.conv_head3_1(dec_3)
        dec_2 = self.conv2_1(ops.concat_channels(enc_2, self.unpool(dec_3, i2)))
        dec_2 = self.SPADE2_1(dec_2, dec_2)
        dec_1 = self.conv1_1(ops.concat_channels(enc_1, self.up(dec_2)))
        dec_1 = self.SPADE1_1(dec_1, dec_1)
        dec_0 = self.conv0_1(ops.concat_channels(enc_0, self.up(dec_1)))
        dec_0 = self.SPADE0_1(dec_0, dec_0)
        nc = self.final.out_channels
        return ops.to_nchw_f32(self.final(dec_0, cout_store=ops.thin_pad(nc)), channels=nc)
