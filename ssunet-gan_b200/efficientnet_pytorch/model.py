"""EfficientNet encoder (reference: efficientnet_pytorch/model.py) on hand-written sm_100a kernels.

Module tree, parameter registration order and state_dict keys follow the reference (`_conv_stem`, `_bn0`,
`_blocks.N._expand_conv / _bn0 / _depthwise_conv / _bn1 / _se_reduce / _se_expand / _project_conv / _bn2`,
`_conv_head`, `_bn1`, `_fc`).  Data path per MBConv block (model.py:64-94):

    1x1 expand (tcgen05 implicit GEMM, BN sums from its epilogue) -> BN-apply -> swish
    depthwise k x k, stride 1/2, TF-same padding (csrc/mbconv.cu) -> BN (stats pass + apply) -> swish
    squeeze-excite: plane sums -> gate MLP (one block per sample) -> scale      (3 launches, 3 more in backward)
    1x1 project (tcgen05) -> BN-apply fused with the identity skip

Building blocks return activations in the package's internal storage (logical NCHW shape, NHWC memory, compute dtype);
`EfficientNet.forward` returns fp32 logits like the reference.
"""
import torch
from torch import nn

from .. import ops
from .._lib import ACT_NONE
from ..nn_layers import BatchNorm2d, Linear
from .utils import (
    MemoryEfficientSwish,
    Swish,
    drop_connect,
    efficientnet_params,
    get_model_params,
    get_same_padding_conv2d,
    load_pretrained_weights,
    round_filters,
    round_repeats,
)


class MBConvBlock(nn.Module):
    """Mobile inverted-residual bottleneck with squeeze-and-excitation (model.py:18-99)."""

    def __init__(self, block_args, global_params):
        super().__init__()
        self._block_args = block_args
        self._bn_mom = 1 - global_params.batch_norm_momentum
        self._bn_eps = global_params.batch_norm_epsilon
        self.has_se = (block_args.se_ratio is not None) and (0 < block_args.se_ratio <= 1)
        self.id_skip = block_args.id_skip
        Conv2d = get_same_padding_conv2d(image_size=global_params.image_size)
        a = block_args
        inp = a.input_filters
        oup = a.input_filters * a.expand_ratio
        if a.expand_ratio != 1:
            self._expand_conv = Conv2d(in_channels=inp, out_channels=oup, kernel_size=1, bias=False)
            self._bn0 = BatchNorm2d(num_features=oup, momentum=self._bn_mom, eps=self._bn_eps)
        self._depthwise_conv = Conv2d(in_channels=oup, out_channels=oup, groups=oup, kernel_size=a.kernel_size, stride=a.stride,
                                      bias=False)
        self._bn1 = BatchNorm2d(num_features=oup, momentum=self._bn_mom, eps=self._bn_eps)
        if self.has_se:
            squeezed = max(1, int(a.input_filters * a.se_ratio))
            self._se_reduce = Conv2d(in_channels=oup, out_channels=squeezed, kernel_size=1)
            self._se_expand = Conv2d(in_channels=squeezed, out_channels=oup, kernel_size=1)
        self._project_conv = Conv2d(in_channels=oup, out_channels=a.output_filters, kernel_size=1, bias=False)
        self._bn2 = BatchNorm2d(num_features=a.output_filters, momentum=self._bn_mom, eps=self._bn_eps)
        self._swish = MemoryEfficientSwish()

    def forward(self, inputs, drop_connect_rate=None):
        a = self._block_args
        inputs = ops.to_nhwc(inputs)
        x = inputs
        if a.expand_ratio != 1:
            y, sums = self._expand_conv(x, want_stats=self.training)
            x = self._swish(self._bn0(y, sums=sums))
        x = self._swish(self._bn1(self._depthwise_conv(x)))
        if self.has_se:
            x = ops.squeeze_excite(x, self._se_reduce.weight, self._se_reduce.bias, self._se_expand.weight, self._se_expand.bias)
        y, sums = self._project_conv(x, want_stats=self.training)
        # the reference compares the raw field (model.py:90): a stage's first block carries stride=[s] (a list), which never
        # equals 1, so only the repeated blocks (stride=1 after _replace, model.py:178) take the identity skip
        skip = self.id_skip and a.stride == 1 and a.input_filters == a.output_filters
        dropping = bool(drop_connect_rate) and self.training
        if skip and not dropping:
            return self._bn2(y, residual=inputs, act=ACT_NONE, sums=sums)       # BN-apply + skip add in one pass
        x = self._bn2(y, sums=sums)
        if skip:
            x = ops.add(drop_connect(x, p=drop_connect_rate, training=self.training), inputs)
        return x

    def set_swish(self, memory_efficient=True):
        self._swish = MemoryEfficientSwish() if memory_efficient else Swish()


class EfficientNet(nn.Module):
    """model.py:136-260."""

    def __init__(self, blocks_args=None, global_params=None):
        super().__init__()
        assert isinstance(blocks_args, list), "blocks_args should be a list"
        assert len(blocks_args) > 0, "block args must be greater than 0"
        self._global_params = global_params
        self._blocks_args = blocks_args
        Conv2d = get_same_padding_conv2d(image_size=global_params.image_size)
        bn_mom = 1 - global_params.batch_norm_momentum
        bn_eps = global_params.batch_norm_epsilon

        stem_out = round_filters(32, global_params)
        self._conv_stem = Conv2d(3, stem_out, kernel_size=3, stride=2, bias=False)
        self._bn0 = BatchNorm2d(num_features=stem_out, momentum=bn_mom, eps=bn_eps)

        self._blocks = nn.ModuleList([])
        for args in blocks_args:
            args = args._replace(input_filters=round_filters(args.input_filters, global_params),
                                 output_filters=round_filters(args.output_filters, global_params),
                                 num_repeat=round_repeats(args.num_repeat, global_params))
            self._blocks.append(MBConvBlock(args, global_params))          # first block of a stage: stride + widening
            if args.num_repeat > 1:
                args = args._replace(input_filters=args.output_filters, stride=1)
            for _ in range(args.num_repeat - 1):
                self._blocks.append(MBConvBlock(args, global_params))

        head_in = args.output_filters
        head_out = round_filters(1280, global_params)
        self._conv_head = Conv2d(head_in, head_out, kernel_size=1, bias=False)
        self._bn1 = BatchNorm2d(num_features=head_out, momentum=bn_mom, eps=bn_eps)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(global_params.dropout_rate)
        self._fc = Linear(head_out, global_params.num_classes)
        self._swish = MemoryEfficientSwish()

    def set_swish(self, memory_efficient=True):
        self._swish = MemoryEfficientSwish() if memory_efficient else Swish()
        for block in self._blocks:
            block.set_swish(memory_efficient)

    def extract_features(self, inputs):
        """Output of the last convolution (model.py:202-218), internal storage."""
        x = self._swish(self._bn0(self._conv_stem(ops.to_nhwc(inputs))))
        for idx, block in enumerate(self._blocks):
            rate = self._global_params.drop_connect_rate
            if rate:
                rate *= float(idx) / len(self._blocks)
            x = block(x, drop_connect_rate=rate)
        y, sums = self._conv_head(x, want_stats=self.training)
        return self._swish(self._bn1(y, sums=sums))

    def forward(self, inputs):
        x = self.extract_features(inputs)
        n, c = x.shape[0], x.shape[1]
        x = ops.adaptive_avg_pool_flat(x, 1, 1)                            # [N, C]
        if self.training and self._dropout.p > 0:
            keep = 1.0 - self._dropout.p
            mask = (torch.rand((n, c), device=x.device) < keep).float() / keep
            x = ops._PlaneScale.apply(x.reshape(n, c, 1, 1), mask).reshape(n, c)
        return self._fc(x).float()

    @classmethod
    def from_name(cls, model_name, override_params=None):
        cls._check_model_name_is_valid(model_name)
        blocks_args, global_params = get_model_params(model_name, override_params)
        return cls(blocks_args, global_params)

    @classmethod
    def from_pretrained(cls, model_name, base_path, advprop=False, num_classes=1000, in_channels=3):
        model = cls.from_name(model_name, override_params={"num_classes": num_classes})
        load_pretrained_weights(model, model_name, base_path, load_fc=(num_classes == 1000), advprop=advprop)
        if in_channels != 3:
            Conv2d = get_same_padding_conv2d(image_size=model._global_params.image_size)
            model._conv_stem = Conv2d(in_channels, round_filters(32, model._global_params), kernel_size=3, stride=2, bias=False)
        return model

    @classmethod
    def get_image_size(cls, model_name):
        cls._check_model_name_is_valid(model_name)
        return efficientnet_params(model_name)[2]

    @classmethod
    def _check_model_name_is_valid(cls, model_name):
        valid = ["efficientnet-b" + str(i) for i in range(9)]
        if model_name not in valid:
            raise ValueError("model_name should be one of: " + ", ".join(valid))
