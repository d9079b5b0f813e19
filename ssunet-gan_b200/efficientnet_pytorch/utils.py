"""Model-definition helpers of the EfficientNet encoder (reference: efficientnet_pytorch/utils.py).

Same public names and meaning as the reference so `EfficientNet.from_name(...)` builds modules with identical
state_dict keys and shapes: the two namedtuples, filter/repeat rounding (utils.py:56-78), the block-string
notation (utils.py:172-251), the "same"-padding convolutions (utils.py:96-141), swish (utils.py:36-53) and
drop_connect (utils.py:81-93).  Forward passes run on csrc/mbconv.cu + the tcgen05 1x1 convolutions.
"""
import collections
import math
import os
import re
from functools import partial

import torch
from torch import nn

from .. import ops
from .._lib import ACT_NONE

GlobalParams = collections.namedtuple("GlobalParams", [
    "batch_norm_momentum", "batch_norm_epsilon", "dropout_rate", "num_classes", "width_coefficient",
    "depth_coefficient", "depth_divisor", "min_depth", "drop_connect_rate", "image_size"])
BlockArgs = collections.namedtuple("BlockArgs", [
    "kernel_size", "num_repeat", "input_filters", "output_filters", "expand_ratio", "id_skip", "stride", "se_ratio"])
GlobalParams.__new__.__defaults__ = (None,) * len(GlobalParams._fields)
BlockArgs.__new__.__defaults__ = (None,) * len(BlockArgs._fields)

# name -> (width multiplier, depth multiplier, native resolution, head dropout)   (utils.py:153-169)
_COEFFS = {
    "efficientnet-b0": (1.0, 1.0, 224, 0.2), "efficientnet-b1": (1.0, 1.1, 240, 0.2),
    "efficientnet-b2": (1.1, 1.2, 260, 0.3), "efficientnet-b3": (1.2, 1.4, 300, 0.3),
    "efficientnet-b4": (1.4, 1.8, 380, 0.4), "efficientnet-b5": (1.6, 2.2, 456, 0.4),
    "efficientnet-b6": (1.8, 2.6, 528, 0.5), "efficientnet-b7": (2.0, 3.1, 600, 0.5),
    "efficientnet-b8": (2.2, 3.6, 672, 0.5), "efficientnet-l2": (4.3, 5.3, 800, 0.5),
}
# the seven stages of the base network (utils.py:258-263)
_STAGES = ["r1_k3_s11_e1_i32_o16_se0.25", "r2_k3_s22_e6_i16_o24_se0.25", "r2_k5_s22_e6_i24_o40_se0.25",
           "r3_k3_s22_e6_i40_o80_se0.25", "r3_k5_s11_e6_i80_o112_se0.25", "r4_k5_s22_e6_i112_o192_se0.25",
           "r1_k3_s11_e6_i192_o320_se0.25"]
# checkpoint file names the reference looks for under `base_path` (utils.py:319-343)
model_map = {"efficientnet-b0": "efficientnet-b0-355c32eb.pth", "efficientnet-b1": "efficientnet-b1-f1951068.pth",
             "efficientnet-b2": "efficientnet-b2-8bb594d6.pth", "efficientnet-b3": "efficientnet-b3-5fb5a3c3.pth",
             "efficientnet-b4": "efficientnet-b4-6ed6700e.pth", "efficientnet-b5": "efficientnet-b5-b6417697.pth",
             "efficientnet-b6": "efficientnet-b6-c76e70fd.pth", "efficientnet-b7": "efficientnet-b7-dcc49843.pth"}
model_map_advprop = {"efficientnet-b0": "adv-efficientnet-b0-b64d5a18.pth", "efficientnet-b1": "adv-efficientnet-b1-0f3ce85a.pth",
                     "efficientnet-b2": "adv-efficientnet-b2-6e9d97e5.pth", "efficientnet-b3": "adv-efficientnet-b3-cdd7c0f4.pth",
                     "efficientnet-b4": "adv-efficientnet-b4-44fb3a87.pth", "efficientnet-b5": "adv-efficientnet-b5-86493f6b.pth",
                     "efficientnet-b6": "adv-efficientnet-b6-ac80338e.pth", "efficientnet-b7": "adv-efficientnet-b7-4652b6dd.pth",
                     "efficientnet-b8": "adv-efficientnet-b8-22a8fe65.pth"}


class MemoryEfficientSwish(nn.Module):
    """x * sigmoid(x); the backward recomputes the sigmoid from the saved input (utils.py:36-53)."""

    def forward(self, x):
        return ops.swish(x)


class Swish(MemoryEfficientSwish):
    pass


def round_filters(filters, global_params):
    """Width-scaled channel count snapped to `depth_divisor`, never more than 10 % below the scaled value."""
    mult = global_params.width_coefficient
    if not mult:
        return filters
    div = global_params.depth_divisor
    floor = global_params.min_depth or div
    scaled = filters * mult
    snapped = max(floor, int(scaled + div / 2) // div * div)
    if snapped < 0.9 * scaled:
        snapped += div
    return int(snapped)


def round_repeats(repeats, global_params):
    mult = global_params.depth_coefficient
    return repeats if not mult else int(math.ceil(mult * repeats))


def drop_connect(inputs, p, training):
    """Stochastic depth: whole samples are zeroed with probability p and the rest rescaled by 1 / (1 - p)."""
    if not training:
        return inputs
    keep = 1.0 - p
    mask = torch.floor(keep + torch.rand([inputs.shape[0]], dtype=torch.float32, device=inputs.device))
    return ops.sample_scale(inputs, mask / keep)


def _same_pad(size, k, stride, dilation=1):
    """Total TF-"same" padding for one spatial dim of extent `size`."""
    out = math.ceil(size / stride)
    return max((out - 1) * stride + (k - 1) * dilation + 1 - size, 0)


class _SamePadConv2d(nn.Conv2d):
    """Shared forward of the two "same"-padding convolutions: depthwise layers run the NHWC depthwise kernel with the
    leading pads folded in; dense layers (the stem and the 1x1s) run the implicit-GEMM kernels after an explicit pad."""

    def _pads(self, x):
        raise NotImplementedError

    def forward(self, x, want_stats=None):
        assert self.dilation[0] == 1 and self.dilation[1] == 1 and self.kernel_size[0] == self.kernel_size[1]
        assert self.stride[0] == self.stride[1]
        pad_h, pad_w = self._pads(x)
        pt, pl = pad_h // 2, pad_w // 2
        k, s = self.kernel_size[0], self.stride[0]
        x = ops.to_nhwc(x)
        if self.groups > 1:
            assert self.groups == self.in_channels == self.out_channels, "only depthwise grouping is used by EfficientNet"
            oh = (x.shape[2] + pad_h - k) // s + 1
            ow = (x.shape[3] + pad_w - k) // s + 1
            y = ops.depthwise_conv2d(x, self.weight, self.bias, s, pt, pl, (oh, ow))
            return (y, None) if want_stats is not None else y
        if pad_h or pad_w:
            x = ops.zero_pad2d(x, pl, pad_w - pl, pt, pad_h - pt)
        return ops.conv2d(x, self.weight, self.bias, s, 0, ACT_NONE, 0.0, None, want_stats)


class Conv2dDynamicSamePadding(_SamePadConv2d):
    """Padding computed from the size of each input (utils.py:106-124)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, groups=1, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride, 0, dilation, groups, bias)
        self.stride = self.stride if len(self.stride) == 2 else [self.stride[0]] * 2

    def _pads(self, x):
        k = self.kernel_size[0]
        return _same_pad(x.shape[-2], k, self.stride[0]), _same_pad(x.shape[-1], k, self.stride[1])


class Identity(nn.Module):
    def forward(self, input):
        return input


class Conv2dStaticSamePadding(_SamePadConv2d):
    """Padding fixed at construction from the network's nominal `image_size` -- NOT from the feature map the layer
    actually sees (utils.py:127-146); the quirk is kept because it decides where the zero rows go."""

    def __init__(self, in_channels, out_channels, kernel_size, image_size=None, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, **kwargs)
        self.stride = self.stride if len(self.stride) == 2 else [self.stride[0]] * 2
        assert image_size is not None
        ih, iw = image_size if type(image_size) == list else [image_size, image_size]
        k = self.kernel_size[0]
        self._pad_hw = (_same_pad(ih, k, self.stride[0], self.dilation[0]), _same_pad(iw, k, self.stride[1], self.dilation[1]))
        ph, pw = self._pad_hw
        # parameter-free child kept so that module trees print / traverse like the reference's
        self.static_padding = nn.ZeroPad2d((pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)) if (ph > 0 or pw > 0) else Identity()

    def _pads(self, x):
        return self._pad_hw


def get_same_padding_conv2d(image_size=None):
    return Conv2dDynamicSamePadding if image_size is None else partial(Conv2dStaticSamePadding, image_size=image_size)


def efficientnet_params(model_name):
    return _COEFFS[model_name]


class BlockDecoder(object):
    """'r2_k3_s22_e6_i16_o24_se0.25' <-> BlockArgs."""

    @staticmethod
    def _decode_block_string(block_string):
        assert isinstance(block_string, str)
        opt = {}
        for tok in block_string.split("_"):
            m = re.match(r"([a-z]+)(\d.*)$", tok)
            if m:
                opt[m.group(1)] = m.group(2)
        s = opt["s"]
        assert len(s) == 1 or (len(s) == 2 and s[0] == s[1]), "stride must be isotropic"
        return BlockArgs(kernel_size=int(opt["k"]), num_repeat=int(opt["r"]), input_filters=int(opt["i"]),
                         output_filters=int(opt["o"]), expand_ratio=int(opt["e"]), id_skip=("noskip" not in block_string),
                         se_ratio=float(opt["se"]) if "se" in opt else None, stride=[int(s[0])])

    @staticmethod
    def _encode_block_string(block):
        s = block.stride if isinstance(block.stride, (list, tuple)) else [block.stride]
        parts = ["r%d" % block.num_repeat, "k%d" % block.kernel_size, "s%d%d" % (s[0], s[-1]), "e%s" % block.expand_ratio,
                 "i%d" % block.input_filters, "o%d" % block.output_filters]
        if block.se_ratio is not None and 0 < block.se_ratio <= 1:
            parts.append("se%s" % block.se_ratio)
        if block.id_skip is False:
            parts.append("noskip")
        return "_".join(parts)

    @staticmethod
    def decode(string_list):
        assert isinstance(string_list, list)
        return [BlockDecoder._decode_block_string(s) for s in string_list]

    @staticmethod
    def encode(blocks_args):
        return [BlockDecoder._encode_block_string(b) for b in blocks_args]


def efficientnet(width_coefficient=None, depth_coefficient=None, dropout_rate=0.2, drop_connect_rate=0.2, image_size=None,
                 num_classes=1000):
    gp = GlobalParams(batch_norm_momentum=0.99, batch_norm_epsilon=1e-3, dropout_rate=dropout_rate,
                      drop_connect_rate=drop_connect_rate, num_classes=num_classes, width_coefficient=width_coefficient,
                      depth_coefficient=depth_coefficient, depth_divisor=8, min_depth=None, image_size=image_size)
    return BlockDecoder.decode(list(_STAGES)), gp


def get_model_params(model_name, override_params):
    if not model_name.startswith("efficientnet"):
        raise NotImplementedError("model name is not pre-defined: %s" % model_name)
    w, d, res, p = efficientnet_params(model_name)
    blocks_args, gp = efficientnet(width_coefficient=w, depth_coefficient=d, dropout_rate=p, image_size=res)
    if override_params:
        gp = gp._replace(**override_params)      # unknown fields raise ValueError, as in the reference
    return blocks_args, gp


def load_pretrained_weights(model, model_name, base_path, load_fc=True, advprop=False):
    """Loads `<base_path>/<checkpoint name>` (utils.py:346-364); there is no download path."""
    path = os.path.join(base_path, (model_map_advprop if advprop else model_map)[model_name])
    print("Pretrained Model Path : ,", path)
    state_dict = torch.load(path)
    if load_fc:
        model.load_state_dict(state_dict)
    else:
        state_dict.pop("_fc.weight")
        state_dict.pop("_fc.bias")
        res = model.load_state_dict(state_dict, strict=False)
        assert set(res.missing_keys) == set(["_fc.weight", "_fc.bias"]), "issue loading pretrained weights"
    print("Loaded pretrained weights for {}".format(model_name))
