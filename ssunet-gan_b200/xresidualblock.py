"""xResidualBlock (reference: xresidualblock.py:4-33) on the hand-written kernels.

    x1 = conv_kxk(x) + b                     implicit GEMM (tcgen05 in bf16 mode)
    z  = BN(dw9x9(relu(BN(x1))))             BN-apply+ReLU fused; depthwise 9x9 in csrc/mbconv.cu
    y  = x1 * exp(-z^2)                      one elementwise kernel (Gaussian gate)
    out = BN(conv3x3(y) + b) + x             BN-apply fused with the identity skip

state_dict keys follow the reference's nn.Sequential indices: md.features.0, md.module.{0,2,3}, conv2, bn1.
"""
from torch import nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .nn_layers import BatchNorm2d, Conv2d, ReLU


class Gaussian(nn.Module):
    def forward(self, input):
        import torch
        z = ops.to_nhwc(input)
        return ops.gauss_gate(torch.ones_like(z), z)


class _DepthwiseConv2d(nn.Conv2d):
    def forward(self, x):
        assert self.groups == self.in_channels == self.out_channels and self.stride[0] == 1
        return ops.depthwise_conv2d(x, self.weight, self.bias, 1, self.padding[0], self.padding[1])


class Modulecell(nn.Module):
    def __init__(self, in_channels=1, out_channels=64, kernel_size=3, skernel_size=9):
        super().__init__()
        self.features = nn.Sequential(
            Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=((kernel_size - 1) // 2), bias=True))
        self.module = nn.Sequential(
            BatchNorm2d(out_channels),
            ReLU(),
            _DepthwiseConv2d(out_channels, out_channels, kernel_size=skernel_size, stride=1, padding=((skernel_size - 1) // 2),
                             groups=out_channels),
            BatchNorm2d(out_channels),
            Gaussian())

    def forward(self, x):
        x1, sums = self.features[0](x, want_stats=self.training)
        t = self.module[0](x1, act=ACT_RELU, sums=sums)
        z = self.module[3](self.module[2](t))
        return ops.gauss_gate(x1, z)


class xResidualBlock(nn.Module):
    def __init__(self, in_channels=64, planes=64, kernel_size=3, s=1):
        super().__init__()
        self.md = Modulecell(in_channels, planes, kernel_size)
        self.conv2 = Conv2d(planes, planes, kernel_size, stride=s, padding=1)
        self.bn1 = BatchNorm2d(planes)

    def forward(self, x):
        x = ops.to_nhwc(x)
        y = self.md(x)
        c, sums = self.conv2(y, want_stats=self.training)
        return self.bn1(c, residual=x, act=ACT_NONE, sums=sums)
