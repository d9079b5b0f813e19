"""CPU oracle for the rest of the reference's model zoo (TEST INFRASTRUCTURE ONLY).

Functional restatement, in plain CPU PyTorch fp32 on flat ``state_dict`` mappings, of the seven
`archs.__all__` networks other than UNet_R_SS_v2 (which lives in ssunet_oracle.py) plus ProgUNet.
Only tests/ may import this file.  Pinned against the unmodified reference by
oracle/make_golden_archs.py -> tests/golden/archs_*.npz / archs_layout.json, checked by
tests/test_oracle_golden.py::test_arch_zoo_oracle_matches_reference.

All file:line citations are relative to /root/reference/scripts/.
"""
from __future__ import annotations

import json
import os

import torch
import torch.nn.functional as F

import ssunet_oracle as O

ARCHS = ("UNet", "NestedUNet", "SSUNet", "UNet_ori", "UNet_B_SS", "AttUNet", "UNet_R_SS")   # archs.py:8 minus UNet_R_SS_v2
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def arch_spec(name, deep_supervision=False):
    """[(key, shape)] of `archs.<name>(3, 3, deep_supervision).state_dict()` in registration order, as recorded from the
    reference by make_golden_archs.py."""
    lay = json.load(open(os.path.join(GOLDEN, "archs_layout.json")))
    return [(k, tuple(s)) for k, s in lay[name + ("_ds" if deep_supervision else "")]]


def _conv_bn(sd, pc, pb, x, training, stride=1, pad=0):
    y = F.conv2d(x, sd[pc + ".weight"], sd.get(pc + ".bias"), stride, pad)
    return O.batch_norm(sd, pb, y, training)


def vgg_block(sd, p, x, training=True):
    """archs.py:103-112."""
    out = F.relu(_conv_bn(sd, p + ".conv1", p + ".bn1", x, training, pad=1))
    return F.relu(_conv_bn(sd, p + ".conv2", p + ".bn2", out, training, pad=1))


def bottleneck(sd, p, x, training=True):
    """archs.py:263-269."""
    out = F.relu(_conv_bn(sd, p + ".conv1", p + ".bn1", x, training))
    out = F.relu(_conv_bn(sd, p + ".conv2", p + ".bn2", out, training, pad=1))
    out = _conv_bn(sd, p + ".conv3", p + ".bn3", out, training)
    if p + ".shortcut.0.weight" in sd:
        sc = _conv_bn(sd, p + ".shortcut.0", p + ".shortcut.1", x, training)
    else:
        sc = x
    return F.relu(out + sc)


def conv_block(sd, p, x, training=True):
    """archs.py:831-846."""
    out = F.relu(_conv_bn(sd, p + ".conv.0", p + ".conv.1", x, training, pad=1))
    return F.relu(_conv_bn(sd, p + ".conv.3", p + ".conv.4", out, training, pad=1))


def up_conv(sd, p, x, training=True):
    """archs.py:848-861: nn.Upsample(scale_factor=2) is nearest-neighbour."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return F.relu(_conv_bn(sd, p + ".up.1", p + ".up.2", x, training, pad=1))


def attention_block(sd, p, g, x, training=True):
    """archs.py:136-142."""
    g1 = _conv_bn(sd, p + ".W_g.0", p + ".W_g.1", g, training)
    x1 = _conv_bn(sd, p + ".W_x.0", p + ".W_x.1", x, training)
    psi = F.relu(g1 + x1)
    psi = torch.sigmoid(_conv_bn(sd, p + ".psi.0", p + ".psi.1", psi, training))
    return x * psi


_BLOCK = {"UNet": vgg_block, "ProgUNet": vgg_block, "SSUNet": vgg_block, "NestedUNet": vgg_block, "UNet_B_SS": bottleneck,
          "UNet_R_SS": lambda sd, p, x, training=True: O.basic_block(sd, p, x, training)}


def _plain(sd, name, x, training):
    """UNet / SSUNet / UNet_B_SS / UNet_R_SS forward wiring (archs.py:375-403,532-555,720-742,815-829)."""
    blk = _BLOCK[name]

    def stage(tag, t):
        t = blk(sd, "conv" + tag, t, training)
        if "SPADE%s.x2map.weight" % tag in sd:
            t = O.spade(sd, "SPADE" + tag, t)
        return t

    pool = lambda t: F.max_pool2d(t, 2, 2)
    x0_0 = stage("0_0", x)
    x1_0 = stage("1_0", pool(x0_0))
    x2_0 = stage("2_0", pool(x1_0))
    x3_0 = stage("3_0", pool(x2_0))
    x4_0 = stage("4_0", pool(x3_0))
    if "conv5_0.conv1.weight" in sd:
        x5_0 = stage("5_0", pool(x4_0))
        x4_1 = stage("4_1", torch.cat([x4_0, O._up(x5_0)], 1))
        x3_1 = stage("3_1", torch.cat([x3_0, O._up(x4_1)], 1))
    else:
        x3_1 = stage("3_1", torch.cat([x3_0, O._up(x4_0)], 1))
    x2_2 = stage("2_2", torch.cat([x2_0, O._up(x3_1)], 1))
    x1_3 = stage("1_3", torch.cat([x1_0, O._up(x2_2)], 1))
    x0_4 = stage("0_4", torch.cat([x0_0, O._up(x1_3)], 1))
    return x0_4, x1_3, x2_2, x3_1


def _head(sd, p, x):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"])


def nested_unet(sd, x, training=True):
    """archs.py:904-933."""
    b = lambda tag, t: vgg_block(sd, "conv" + tag, t, training)
    pool = lambda t: F.max_pool2d(t, 2, 2)
    up = O._up
    x0_0 = b("0_0", x)
    x1_0 = b("1_0", pool(x0_0))
    x0_1 = b("0_1", torch.cat([x0_0, up(x1_0)], 1))
    x2_0 = b("2_0", pool(x1_0))
    x1_1 = b("1_1", torch.cat([x1_0, up(x2_0)], 1))
    x0_2 = b("0_2", torch.cat([x0_0, x0_1, up(x1_1)], 1))
    x3_0 = b("3_0", pool(x2_0))
    x2_1 = b("2_1", torch.cat([x2_0, up(x3_0)], 1))
    x1_2 = b("1_2", torch.cat([x1_0, x1_1, up(x2_1)], 1))
    x0_3 = b("0_3", torch.cat([x0_0, x0_1, x0_2, up(x1_2)], 1))
    x4_0 = b("4_0", pool(x3_0))
    x3_1 = b("3_1", torch.cat([x3_0, up(x4_0)], 1))
    x2_2 = b("2_2", torch.cat([x2_0, x2_1, up(x3_1)], 1))
    x1_3 = b("1_3", torch.cat([x1_0, x1_1, x1_2, up(x2_2)], 1))
    x0_4 = b("0_4", torch.cat([x0_0, x0_1, x0_2, x0_3, up(x1_3)], 1))
    if "final1.weight" in sd:
        return [_head(sd, "final1", x0_1), _head(sd, "final2", x0_2), _head(sd, "final3", x0_3), _head(sd, "final4", x0_4)]
    return _head(sd, "final", x0_4)


def att_unet(sd, x, training=True):
    """UNet_ori (archs.py:961-996) and AttUNet (archs.py:303-342): the gates exist iff the state_dict has `Att*` keys."""
    pool = lambda t: F.max_pool2d(t, 2, 2)
    x1 = conv_block(sd, "Conv1", x, training)
    x2 = conv_block(sd, "Conv2", pool(x1), training)
    x3 = conv_block(sd, "Conv3", pool(x2), training)
    x4 = conv_block(sd, "Conv4", pool(x3), training)
    d = conv_block(sd, "Conv5", pool(x4), training)
    for lvl, skip in ((5, x4), (4, x3), (3, x2), (2, x1)):
        d = up_conv(sd, "Up%d" % lvl, d, training)
        if "Att%d.psi.0.weight" % lvl in sd:
            skip = attention_block(sd, "Att%d" % lvl, d, skip, training)
        d = conv_block(sd, "Up_conv%d" % lvl, torch.cat((skip, d), 1), training)
    return _head(sd, "Conv_1x1", d)


def arch_forward(name, sd, x, training=True):
    """Logits (or the list of logits under deep supervision / ProgUNet) of `archs.<name>` on NCHW fp32 `x`."""
    if name == "NestedUNet":
        return nested_unet(sd, x, training)
    if name in ("UNet_ori", "AttUNet"):
        return att_unet(sd, x, training)
    outs = _plain(sd, name, x, training)
    if name == "ProgUNet":
        return [_head(sd, "final%d" % i, o) for i, o in enumerate(outs)]
    return _head(sd, "final", outs[0])


def supervised_loss(outputs, target):
    """train.py:85-108: BCEDiceLoss, averaged over the outputs under deep supervision."""
    if isinstance(outputs, (list, tuple)):
        loss = 0
        for o in outputs:
            loss = loss + O.bce_dice_loss(o, target)
        return loss / len(outputs)
    return O.bce_dice_loss(outputs, target)


def supervised_train_step(name, sd, opt, x, target, clip, num_classes=3):
    """One iteration of train.py:85-116 on the flat state_dict: forward, loss, metrics, the in-place WEIGHT clamp
    (:111-112, after the forward and before the backward: autograd's saved parameters alias the clamped storage, so the
    backward pass reads clamped weights), backward, Adam.  ``opt`` is an ssunet_oracle.AdamState."""
    O._leafify(sd)
    if name == "UNet_R_SS_v2":
        out = O.unet_r_ss_v2(sd, x, True)
    else:
        out = arch_forward(name, sd, x, True)
    if isinstance(out, list):
        loss = supervised_loss(out, target)
        last = out[-1]
        iou, dice = O.iou_score(last, target), O.dice_coef(last, target)
    else:
        out = torch.where(torch.isnan(out), torch.zeros_like(out), out)       # :101
        loss = O.bce_dice_loss(out, target)
        iou = O.iou_score(out[:, 1:num_classes].detach().clone(), target[:, 1:num_classes].clone())
        dice = O.dice_coef(out[:, 1:num_classes].detach().clone(), target[:, 1:num_classes].clone())
        last = out
    keys = O.trainable_keys(sd)
    with torch.no_grad():
        for k in keys:
            sd[k].data.clamp_(-clip, clip)
    grads = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    opt.step(sd, dict(zip(keys, grads)))
    return {"loss": float(loss), "iou": iou, "dice": dice, "logits": last.detach()}


# --------------------------------------------------------------------------------------
# data feed -- numpy restatement
# --------------------------------------------------------------------------------------
def feed_normalize(img_u8, mean, std, max_pixel_value=255.0):
    """albumentations.augmentations.functional.normalize as published (albumentations is NOT installed here, so this
    restatement is unpinned; the reference calls it through transforms.Normalize, train_seg_gan.py:376,381):
    mean *= max_pixel; std *= max_pixel; denominator = reciprocal(std); img = (float32(img) - mean) * denominator."""
    import numpy as np
    mean = np.array(mean, dtype=np.float32) * np.float32(max_pixel_value)
    std = np.array(std, dtype=np.float32) * np.float32(max_pixel_value)
    den = np.reciprocal(std, dtype=np.float32)
    img = img_u8.astype(np.float32)
    img -= mean
    img *= den
    return img


def feed_image(img_u8_nhwc, mean, std, flip_cv2_codes=None):
    """Normalize -> Flip (cv2.flip semantics, as albumentations' Flip applies them) -> `img.astype('float32')`,
    `transpose(2, 0, 1)` (dataset.py:137-138).  Normalising before or after the flip gives identical values."""
    import numpy as np
    out = []
    for i, im in enumerate(img_u8_nhwc):
        f = feed_normalize(im, mean, std)
        if flip_cv2_codes is not None and flip_cv2_codes[i] is not None:
            d = flip_cv2_codes[i]
            f = f[::-1] if d == 0 else (f[:, ::-1] if d == 1 else f[::-1, ::-1])
        out.append(np.ascontiguousarray(f.transpose(2, 0, 1)))
    return np.stack(out)


def feed_mask(mask_u8_nhwc, flip_cv2_codes=None):
    """dataset.py:126-140 for the multi-class layout: `(png.astype('float32') / 255.0).astype('uint8')`, stacked, flipped
    with the image, `1.0 * mask.astype('float32')`, CHW."""
    import numpy as np
    out = []
    for i, m in enumerate(mask_u8_nhwc):
        b = (m.astype("float32") / 255.0).astype("uint8")
        if flip_cv2_codes is not None and flip_cv2_codes[i] is not None:
            d = flip_cv2_codes[i]
            b = b[::-1] if d == 0 else (b[:, ::-1] if d == 1 else b[::-1, ::-1])
        out.append(np.ascontiguousarray((1.0 * b.astype("float32")).transpose(2, 0, 1)))
    return np.stack(out)


# --------------------------------------------------------------------------------------
# cv2.resize(uint8, INTER_LINEAR) restated (the resize bridge of tiled inference:
# aerial_image_segmentation_api.py:151-152 upsamples every uint8 mask patch to patch_size, :361 shrinks every image
# patch to the network's input size).  Pinned against cv2 itself by tests/test_oracle_golden.py.
# Algorithm = OpenCV's resizeGeneric_ for 8U: 11-bit fixed-point coefficients, horizontal pass in int32, vertical pass
# ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2; exact 2x shrinking is rerouted by cv::resize to the
# INTER_AREA fast path (a + b + c + d + 2) >> 2.
# --------------------------------------------------------------------------------------
def cv2_linear_tables(src, dst):
    """(ofs[dst], coef[dst, 2]) of OpenCV's linear resize along one axis of length `src` -> `dst`: source index of the
    left tap and the two 11-bit coefficients (x-axis convention: taps clamped by zeroing the fraction at the borders)."""
    import numpy as np
    scale = 1.0 / (float(dst) / float(src))                    # double, as cv::resize computes it
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    f[lo], s[lo] = 0.0, 0
    hi = s >= src - 1
    f[hi], s[hi] = 0.0, src - 1
    coef = np.stack([np.float32(1.0) - f, f], 1) * np.float32(2048.0)
    return s.astype(np.int32), np.clip(np.rint(coef), -32768, 32767).astype(np.int32)


def cv2_resize_linear_u8(img, dsize):
    """cv2.resize(img, dsize) for uint8 `img` [H, W] or [H, W, C] with the default INTER_LINEAR; dsize = (width, height)."""
    import numpy as np
    ow, oh = dsize
    a = img if img.ndim == 3 else img[..., None]
    h, w, _ = a.shape
    if w == 2 * ow and h == 2 * oh:                            # cv::resize: INTER_LINEAR at exactly 1/2 -> INTER_AREA fast path
        s = a.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
    else:
        xo, xa = cv2_linear_tables(w, ow)
        # the y axis keeps its fraction at the borders and clamps the two ROW indices instead (resizeGeneric_Invoker)
        scale = 1.0 / (float(oh) / float(h))
        fy = ((np.arange(oh, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        sy = np.floor(fy).astype(np.int64)
        fy = (fy - sy.astype(np.float32)).astype(np.float32)
        yb = np.clip(np.rint(np.stack([np.float32(1.0) - fy, fy], 1) * np.float32(2048.0)), -32768, 32767).astype(np.int64)
        y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
        s = a.astype(np.int64)
        x1 = np.minimum(xo + 1, w - 1)
        rows = s[:, xo] * xa[:, 0][None, :, None] + s[:, x1] * xa[:, 1][None, :, None]          # [H, ow, C] int
        r0, r1 = rows[y0], rows[y1]
        out = (((yb[:, 0][:, None, None] * (r0 >> 4)) >> 16) + ((yb[:, 1][:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if img.ndim == 3 else out[..., 0]


def tile_merge_resized(img_h, img_w, masks, p_size, num_classes, overlap):
    """patch_merge (aerial_image_segmentation_api.py:129-217) for maps of ANY size: each (mask * 255).astype('uint8') is
    resized to (p_size, p_size) with cv2.resize's arithmetic (:150-152) before the 127 threshold."""
    import numpy as np
    wins = O.tile_windows(img_h, img_w, p_size, overlap)
    out = []
    for c in range(num_classes):
        merged = np.zeros((img_h, img_w))
        div = np.zeros((img_h, img_w))
        for (h1, w1), m in zip(wins, masks):
            u8 = (np.asarray(m[c]) * 255).astype("uint8")
            u8 = cv2_resize_linear_u8(u8, (p_size, p_size))
            merged[h1:h1 + p_size, w1:w1 + p_size] += O._threshold127(u8) / 255.0
            div[h1:h1 + p_size, w1:w1 + p_size] += 1.0
        div[div == 0] = 1.0
        out.append(O._threshold127((np.divide(merged, div) * 255).astype("uint8")))
    return out


def patched_input(img_u8, p_size, size, overlap, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """get_patched_input (:336-373) without file I/O: patch_gen windows, cv2.resize of each patch to (size, size),
    albumentations' default Normalize() (restated, unpinned: see feed_normalize), `/ 255` again (:367), CHW."""
    import numpy as np
    wins = O.tile_windows(img_u8.shape[0], img_u8.shape[1], p_size, overlap)
    out = []
    for h1, w1 in wins:
        patch = cv2_resize_linear_u8(np.ascontiguousarray(img_u8[h1:h1 + p_size, w1:w1 + p_size]), (size, size))
        f = feed_normalize(patch, mean, std)
        f = f.astype("float32") / 255
        out.append(np.ascontiguousarray(f.transpose(2, 0, 1)))
    return np.array(out)
