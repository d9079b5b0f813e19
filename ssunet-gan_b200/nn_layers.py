"""nn.Module building blocks: torch's parameter containers (same init, same state_dict keys) whose
forward runs the hand-written kernels."""
import torch
from torch import nn

from . import ops
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU  # noqa: F401


class Conv2d(nn.Conv2d):
    """nn.Conv2d drop-in (square kernel, symmetric zero padding, dilation 1, groups 1)."""

    def forward(self, x, act=ACT_NONE, slope=0.0, cout_store=None, want_stats=None, dx_sink=None, input_act=None):
        assert self.groups == 1 and self.dilation == (1, 1) and self.padding_mode == "zeros"
        assert self.kernel_size[0] == self.kernel_size[1] and self.stride[0] == self.stride[1]
        assert self.padding[0] == self.padding[1]
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0], act, slope, cout_store, want_stats, dx_sink,
                          input_act)


class Linear(nn.Linear):
    def forward(self, x, act=ACT_NONE, slope=0.0):
        return ops.linear(x, self.weight, self.bias, act, slope)


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d drop-in; `forward(x, residual, act)` additionally fuses the residual add and the
    activation that follow it in BasicBlock / ConvolutionalBlock."""

    sync_group = None      # set by convert_model / SynchronizedBatchNorm2d
    sync_quirk = False

    def forward(self, x, residual=None, act=ACT_NONE, slope=0.0, sums=None):
        training = self.training or (self.running_mean is None)
        # the counter is advanced by the statistics kernel itself (ops.batch_norm(nbt=...)): no separate launch per layer
        nbt = None
        if training and self.track_running_stats and self.num_batches_tracked is not None and not self.sync_quirk:
            nbt = self.num_batches_tracked
        momentum = self.momentum
        if momentum is None:   # cumulative moving average: the factor depends on the counter's host value
            if nbt is not None:
                nbt.add_(1)
                nbt = None
            momentum = 1.0 / float(self.num_batches_tracked) if training else 0.0
        return ops.batch_norm(x, self.weight, self.bias, self.running_mean if self.track_running_stats else None,
                              self.running_var if self.track_running_stats else None, training, momentum, self.eps,
                              residual, act, slope, self.sync_group, self.sync_quirk, sums, nbt)


class ReLU(nn.Module):
    def forward(self, x):
        return ops.relu(x)


class LeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.01):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return ops.leaky_relu(x, self.negative_slope)
