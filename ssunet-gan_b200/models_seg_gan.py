"""Generator / Discriminator of the seg-GAN (reference: models_seg_gan.py:13-64,193-300)."""
import os

import torch
from torch import nn

from . import archs, ops
from ._lib import ACT_LEAKY, ACT_NONE
from .nn_layers import BatchNorm2d, Conv2d, LeakyReLU, Linear


def remove_prefix(state_dict, prefix):
    return {(k.split(prefix, 1)[-1] if k.startswith(prefix) else k): v for k, v in state_dict.items()}


class ConvolutionalBlock(nn.Module):
    """conv (+bias) -> [BN] -> [activation]   (models_seg_gan.py:13-64).  Only the variants the hot path
    instantiates run fused: LeakyReLU(0.2) is folded into the conv epilogue (no BN) or the BN-apply pass."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, batch_norm=False, activation=None):
        super().__init__()
        if activation is not None:
            activation = activation.lower()
            assert activation in {"prelu", "leakyrelu", "tanh"}
        layers = [Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)]
        if batch_norm is True:
            layers.append(BatchNorm2d(num_features=out_channels))
        if activation == "prelu":
            layers.append(nn.PReLU())
        elif activation == "leakyrelu":
            layers.append(LeakyReLU(0.2))
        elif activation == "tanh":
            layers.append(nn.Tanh())
        self.conv_block = nn.Sequential(*layers)
        self._fusable = activation in (None, "leakyrelu")
        self._act = ACT_LEAKY if activation == "leakyrelu" else ACT_NONE
        self._has_bn = batch_norm is True

    def forward(self, input, input_act=None):
        """input_act = (act, slope): `input` is the output of an activation-fused convolution with no BatchNorm (the previous
        block's `fused_output_act`) and feeds nothing but this block; the data gradient then carries that activation's backward."""
        if not self._fusable:
            raise NotImplementedError("only None / LeakyReLU activations are on the seg-GAN path")
        conv = self.conv_block[0]
        if self._has_bn:
            y, sums = conv(input, want_stats=self.training, input_act=input_act)
            return self.conv_block[1](y, act=self._act, slope=0.2, sums=sums)
        return conv(input, act=self._act, slope=0.2, input_act=input_act)

    @property
    def fused_output_act(self):
        """(act, slope) when this block's output comes straight out of the convolution's activation epilogue (no BatchNorm)."""
        return (self._act, 0.2) if (not self._has_bn and self._act != ACT_NONE) else None


class Generator(nn.Module):
    """Thin wrapper: `self.net = archs.__dict__[config['arch']](...)`   (models_seg_gan.py:193-243)."""

    def __init__(self, config):
        super().__init__()
        self.net = archs.__dict__[config["arch"]](config["num_classes"], config["input_channels"], config["deep_supervision"])

    def initialize_with_srresnet(self, model_folder, config):
        model_dict = torch.load(os.path.join(model_folder, "%s/model.pth" % config["name"]))
        if "state_dict" in model_dict.keys():
            model_dict = remove_prefix(model_dict["state_dict"], "module.")
        else:
            model_dict = remove_prefix(model_dict, "module.")
        self.net.load_state_dict(model_dict, strict=False)
        print("\nLoaded weights from pre-trained SS-UNet-R.\n")

    def forward(self, lr_imgs):
        return self.net(lr_imgs)


class Discriminator(nn.Module):
    """SRGAN discriminator: 8 conv blocks, adaptive 6x6 pool, fc1, LeakyReLU, fc2 -> logit (N, 1)
    (models_seg_gan.py:246-300)."""

    def __init__(self, num_classes, kernel_size=3, n_channels=64, n_blocks=8, fc_size=1024):
        super().__init__()
        in_channels = num_classes
        conv_blocks = []
        out_channels = in_channels
        for i in range(n_blocks):
            out_channels = (n_channels if i == 0 else in_channels * 2) if i % 2 == 0 else in_channels
            conv_blocks.append(ConvolutionalBlock(in_channels, out_channels, kernel_size, stride=1 if i % 2 == 0 else 2,
                                                  batch_norm=i != 0, activation="LeakyReLu"))
            in_channels = out_channels
        self.conv_blocks = nn.Sequential(*conv_blocks)
        self.adaptive_pool = nn.AdaptiveAvgPool2d((6, 6))   # kept for state/introspection; the kernel fuses pool+flatten
        self.fc1 = Linear(out_channels * 6 * 6, fc_size)
        self.leaky_relu = LeakyReLU(0.2)
        self.fc2 = Linear(1024, 1)

    def forward(self, imgs):
        out = ops.to_nhwc(imgs, pad_channels=True)
        prev_act = None
        for blk in self.conv_blocks:        # a Sequential chain: every block's output feeds exactly the next block
            out = blk(out, input_act=prev_act)
            prev_act = blk.fused_output_act
        flat = ops.adaptive_avg_pool_flat(out, 6, 6)
        hid = self.fc1(flat, act=ACT_LEAKY, slope=0.2)
        logit = self.fc2(hid)
        return logit.float() if logit.dtype != torch.float32 else logit
