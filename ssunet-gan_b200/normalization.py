"""SPADE block, self-conditioned as the reference uses it (normalization.py:67-122).

forward(x, segmap): segmap -> x2map (C->label_nc) -> mlp_shared (label_nc->h, ReLU) -> gamma, beta
(h->C each) -> x * (1 + gamma) + beta.  The reference constructs `param_free_norm` but never
calls it (normalization.py:110); it is kept so state_dict keys match.
gamma and beta are produced by ONE convolution with 2C output channels (weights concatenated),
and the modulation is one elementwise kernel; neither 1+gamma nor the product is materialised.
"""
import re

import torch
from torch import nn

from . import ops
from ._lib import ACT_RELU
from .batchnorm import SynchronizedBatchNorm2d
from .nn_layers import BatchNorm2d, Conv2d, ReLU


class SPADE(nn.Module):
    def __init__(self, config_text, norm_nc, label_nc, nhidden=64):
        super().__init__()
        assert config_text.startswith("spade")
        parsed = re.search(r"spade(\D+)(\d)x\d", config_text)
        norm_type = str(parsed.group(1))
        ks = int(parsed.group(2))
        if norm_type == "instance":
            self.param_free_norm = nn.InstanceNorm2d(norm_nc, affine=False)
        elif norm_type == "syncbatch":
            self.param_free_norm = SynchronizedBatchNorm2d(norm_nc, affine=False)
        elif norm_type == "batch":
            self.param_free_norm = BatchNorm2d(norm_nc, affine=False)
        else:
            raise ValueError("%s is not a recognized param-free norm type in SPADE" % norm_type)
        nhidden = int(max(nhidden, 4))
        pw = ks // 2
        self.mlp_shared = nn.Sequential(Conv2d(label_nc, nhidden, kernel_size=ks, padding=pw), ReLU())
        self.x2map = Conv2d(norm_nc, label_nc, kernel_size=ks, padding=pw)
        self.mlp_gamma = Conv2d(nhidden, norm_nc, kernel_size=ks, padding=pw)
        self.mlp_beta = Conv2d(nhidden, norm_nc, kernel_size=ks, padding=pw)
        self._pw = pw

    def forward(self, x, segmap):
        x = ops.to_nhwc(x)
        if (segmap is x and self._pw == 1 and x.dtype == torch.bfloat16
                and ops.spade_fused_enabled(x.shape[1], self.x2map.out_channels, self.mlp_shared[0].out_channels)):
            return ops.spade_fused(x, self.x2map, self.mlp_shared[0], self.mlp_gamma, self.mlp_beta)      # opt-in, DESIGN.md §7.1
        # the 3- and h-channel maps are stored channel-padded (ops.thin_pad) so they can feed the TMA/tcgen05 kernels
        # self-conditioned use (segmap is x): x feeds x2map AND the modulation; their two gradient contributions meet in one
        # buffer (ops.GradSink: the x2map data-gradient kernel adds into what the modulation's backward wrote)
        sink = ops.grad_sink_for(x) if segmap is x else None
        seg = self.x2map(segmap, cout_store=ops.thin_pad(self.x2map.out_channels), dx_sink=sink)
        actv = self.mlp_shared[0](seg, act=ACT_RELU, cout_store=ops.thin_pad(self.mlp_shared[0].out_channels))
        w_gb = torch.cat([self.mlp_gamma.weight, self.mlp_beta.weight], 0)
        b_gb = torch.cat([self.mlp_gamma.bias, self.mlp_beta.bias], 0)
        gb = ops.conv2d(actv, w_gb, b_gb, 1, self._pw)
        return ops.spade_modulate(x, gb, dx_sink=sink)
