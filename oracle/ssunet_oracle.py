"""CPU oracle for the seg-GAN training-step hot path (TEST INFRASTRUCTURE ONLY).

This file is a functional restatement, in plain CPU PyTorch fp32, of the arithmetic the
reference performs on the path named by BASELINE.json:north_star.  It is NOT product code:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product (ssunet-gan_b200/) never routes through this file and fails loudly
when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors.  The oracle is
pinned against the reference itself: oracle/make_golden.py imports the unmodified modules from
/root/reference/scripts in the build container, drives them with the portable weights/inputs
defined here and writes tests/golden/*.npz; tests/test_oracle_golden.py checks this file
against those fixtures.

Every function operates on a flat ``state_dict``-style mapping (same keys/shapes as the
reference modules' ``state_dict()``), so key compatibility is checked implicitly.

All file:line citations are relative to /root/reference/scripts/.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

NB_FILTER = (64, 128, 256, 384, 512, 768)  # archs.py:568


# --------------------------------------------------------------------------------------
# state_dict specs (keys, shapes, order) and portable deterministic weights
# --------------------------------------------------------------------------------------
def _bn_keys(prefix, c, affine=True):
    out = []
    if affine:
        out += [(prefix + ".weight", (c,)), (prefix + ".bias", (c,))]
    out += [(prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)),
            (prefix + ".num_batches_tracked", ())]
    return out


def _basic_block_keys(p, cin, c):
    """archs.py:205-219 (BasicBlock.__init__)."""
    k = [(p + ".conv1.weight", (c, cin, 3, 3))]
    k += _bn_keys(p + ".bn1", c)
    k += [(p + ".conv2.weight", (c, c, 3, 3))]
    k += _bn_keys(p + ".bn2", c)
    if cin != c:
        k += [(p + ".shortcut.0.weight", (c, cin, 1, 1))]
    return k


def spade_hidden(c, ss_scale=16):
    """normalization.py:88 with nhidden = C / 16 passed from archs.py:575-613."""
    return int(max(c / ss_scale, 4))


def _spade_keys(p, c, label_nc):
    """normalization.py:67-98 (SPADE.__init__), registration order."""
    h = spade_hidden(c)
    k = _bn_keys(p + ".param_free_norm", c, affine=False)
    k += [(p + ".mlp_shared.0.weight", (h, label_nc, 3, 3)), (p + ".mlp_shared.0.bias", (h,)),
          (p + ".x2map.weight", (label_nc, c, 3, 3)), (p + ".x2map.bias", (label_nc,)),
          (p + ".mlp_gamma.weight", (c, h, 3, 3)), (p + ".mlp_gamma.bias", (c,)),
          (p + ".mlp_beta.weight", (c, h, 3, 3)), (p + ".mlp_beta.bias", (c,))]
    return k


def unet_r_ss_v2_spec(num_classes=3, input_channels=3, prefix=""):
    """Key/shape list of UNet_R_SS_v2.state_dict() in registration order (archs.py:559-617)."""
    f = NB_FILTER
    k = []

    def blk(name, cin, c):
        k.extend(_basic_block_keys(prefix + name, cin, c))

    def spd(name, c):
        k.extend(_spade_keys(prefix + name, c, num_classes))

    blk("conv0_0", input_channels, f[0]); spd("SPADE0_0", f[0])
    blk("conv1_0", f[0], f[1]); spd("SPADE1_0", f[1])
    blk("conv2_0", f[1], f[2]); spd("SPADE2_0", f[2])
    blk("conv3_0", f[2], f[3]); spd("SPADE3_0", f[3])
    blk("conv4_0", f[3], f[4]); spd("SPADE4_0", f[4])
    blk("conv5_0", f[4], f[5]); spd("SPADE5_0", f[5])
    k.append((prefix + "conv_head5_0.weight", (f[4], f[5], 1, 1)))
    blk("conv4_1", f[4] + f[4], f[4]); spd("SPADE4_1", f[4])
    k.append((prefix + "conv_head4_1.weight", (f[3], f[4], 1, 1)))
    blk("conv3_1", f[3] + f[3], f[3]); spd("SPADE3_1", f[3])
    k.append((prefix + "conv_head3_1.weight", (f[2], f[3], 1, 1)))
    blk("conv2_1", f[2] + f[2], f[2]); spd("SPADE2_1", f[2])
    blk("conv1_1", f[1] + f[2], f[1]); spd("SPADE1_1", f[1])
    blk("conv0_1", f[0] + f[1], f[0]); spd("SPADE0_1", f[0])
    k += [(prefix + "final.weight", (num_classes, f[0], 1, 1)), (prefix + "final.bias", (num_classes,))]
    return k


def discriminator_spec(num_classes=3, kernel_size=3, n_channels=64, n_blocks=8, fc_size=1024):
    """Key/shape list of Discriminator.state_dict() (models_seg_gan.py:246-283)."""
    k = []
    cin = num_classes
    cout = cin
    for i in range(n_blocks):
        cout = (n_channels if i == 0 else cin * 2) if i % 2 == 0 else cin
        p = "conv_blocks.%d.conv_block" % i
        k += [(p + ".0.weight", (cout, cin, kernel_size, kernel_size)), (p + ".0.bias", (cout,))]
        if i != 0:
            k += _bn_keys(p + ".1", cout)
        cin = cout
    k += [("fc1.weight", (fc_size, cout * 36)), ("fc1.bias", (fc_size,)),
          ("fc2.weight", (1, 1024)), ("fc2.bias", (1,))]
    return k


def portable_tensor(key, shape, salt=0):
    """Deterministic, machine-independent tensor for ``key`` (CPU generator, one stream per key).

    conv/linear weights ~ N(0, 1/fan_in) * 0.58 (std of the default kaiming_uniform(a=sqrt 5)), BN weight ~ 1 + 0.1 N, biases ~ 0.1 N,
    running_mean ~ 0.1 N, running_var ~ 1 + 0.2 U, num_batches_tracked = 0.
    """
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * salt) & 0x7FFFFFFF)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.int64)
    if leaf == "running_mean":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "running_var":
        return 1.0 + 0.2 * torch.rand(shape, generator=g)
    if len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        return torch.randn(shape, generator=g) * (0.58 / math.sqrt(fan_in))
    if leaf == "weight":  # 1-D weight == BN gamma
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.1 * torch.randn(shape, generator=g)


def portable_state_dict(spec, salt=0):
    return OrderedDict((k, portable_tensor(k, s, salt)) for k, s in spec)


def synthetic_batch(batch, cin, h, w, num_classes=3, seed=1234, blobby=False):
    """input = randn, target = (rand > 0.5); draw order input then target."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cin, h, w, generator=g)
    if not blobby:
        t = (torch.rand(batch, num_classes, h, w, generator=g) > 0.5).float()
    else:
        n = torch.rand(batch, num_classes, h, w, generator=g)
        k = min(31, (min(h, w) // 2) * 2 - 1)
        n = F.avg_pool2d(n, k, stride=1, padding=k // 2)
        t = (n > n.flatten(1).median(dim=1).values.view(-1, 1, 1, 1)).float()
        t[:, 0] = 1.0 - t[:, 1:].amax(dim=1)
    return x, t


# --------------------------------------------------------------------------------------
# primitive layers
# --------------------------------------------------------------------------------------
def batch_norm(sd, p, x, training, eps=1e-5, momentum=0.1, sync_stats=None):
    """nn.BatchNorm2d (train/eval) and SynchronizedBatchNorm2d parallel-mode arithmetic.

    ``sync_stats``: None -> F.batch_norm semantics (batchnorm.py:52-55, also what nn.BatchNorm2d does).
    Otherwise a callable ``(sum, ssum, count) -> (sum, ssum, count)`` performing the cross-replica
    reduction; the arithmetic then follows batchnorm.py:57-80,115-127, including clamp(eps)
    instead of +eps, attribute re-binding of running stats and NO num_batches_tracked update.
    """
    w = sd.get(p + ".weight")
    b = sd.get(p + ".bias")
    if sync_stats is None or not training:
        if training and p + ".num_batches_tracked" in sd:
            sd[p + ".num_batches_tracked"] += 1
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], w, b,
                            training, momentum, eps)
    shape = x.shape
    xf = x.reshape(shape[0], shape[1], -1)
    count = xf.shape[0] * xf.shape[2]
    s = xf.sum(0).sum(-1)               # batchnorm.py:26-28 (_sum_ft)
    ss = (xf ** 2).sum(0).sum(-1)
    s, ss, count = sync_stats(s, ss, count)
    mean = s / count                    # batchnorm.py:118-121
    sumvar = ss - s * mean
    unbias_var = sumvar / (count - 1)
    bias_var = sumvar / count
    with torch.no_grad():               # batchnorm.py:124-125
        sd[p + ".running_mean"] = (1 - momentum) * sd[p + ".running_mean"] + momentum * mean.detach()
        sd[p + ".running_var"] = (1 - momentum) * sd[p + ".running_var"] + momentum * unbias_var.detach()
    inv_std = bias_var.clamp(eps) ** -0.5   # batchnorm.py:127
    if w is not None:                   # batchnorm.py:74-77
        y = (xf - mean.view(1, -1, 1)) * (inv_std * w).view(1, -1, 1) + b.view(1, -1, 1)
    else:
        y = (xf - mean.view(1, -1, 1)) * inv_std.view(1, -1, 1)
    return y.reshape(shape)


def basic_block(sd, p, x, training=True, sync_stats=None):
    """archs.py:229-234: relu(bn2(conv2(relu(bn1(conv1 x)))) + shortcut(x))."""
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1)
    out = F.relu(batch_norm(sd, p + ".bn1", out, training, sync_stats=sync_stats))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
    out = batch_norm(sd, p + ".bn2", out, training, sync_stats=sync_stats)
    if p + ".shortcut.0.weight" in sd:
        sc = F.conv2d(x, sd[p + ".shortcut.0.weight"])
    else:
        sc = x
    return F.relu(out + sc)


def spade(sd, p, x):
    """normalization.py:106-122 with segmap = x; param_free_norm is skipped (:110)."""
    seg = F.conv2d(x, sd[p + ".x2map.weight"], sd[p + ".x2map.bias"], 1, 1)
    actv = F.relu(F.conv2d(seg, sd[p + ".mlp_shared.0.weight"], sd[p + ".mlp_shared.0.bias"], 1, 1))
    gamma = F.conv2d(actv, sd[p + ".mlp_gamma.weight"], sd[p + ".mlp_gamma.bias"], 1, 1)
    beta = F.conv2d(actv, sd[p + ".mlp_beta.weight"], sd[p + ".mlp_beta.bias"], 1, 1)
    return x * (1 + gamma) + beta


def _up(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)  # archs.py:573


def unet_r_ss_v2(sd, x, training=True, sync_stats=None, prefix=""):
    """archs.py:623-671 (UNet_R_SS_v2.forward)."""
    P = prefix

    def stage(name_c, name_s, t):
        t = basic_block(sd, P + name_c, t, training, sync_stats)
        return spade(sd, P + name_s, t)

    enc0 = stage("conv0_0", "SPADE0_0", x)
    p0, _ = F.max_pool2d(enc0, 2, 2, return_indices=True)
    enc1 = stage("conv1_0", "SPADE1_0", p0)
    p1, _ = F.max_pool2d(enc1, 2, 2, return_indices=True)
    enc2 = stage("conv2_0", "SPADE2_0", p1)
    p2, i2 = F.max_pool2d(enc2, 2, 2, return_indices=True)
    enc3 = stage("conv3_0", "SPADE3_0", p2)
    p3, i3 = F.max_pool2d(enc3, 2, 2, return_indices=True)
    enc4 = stage("conv4_0", "SPADE4_0", p3)
    p4, i4 = F.max_pool2d(enc4, 2, 2, return_indices=True)
    enc5 = stage("conv5_0", "SPADE5_0", p4)
    enc5 = F.conv2d(enc5, sd[P + "conv_head5_0.weight"])
    dec4 = stage("conv4_1", "SPADE4_1", torch.cat([enc4, F.max_unpool2d(enc5, i4, 2, 2)], 1))
    dec4 = F.conv2d(dec4, sd[P + "conv_head4_1.weight"])
    dec3 = stage("conv3_1", "SPADE3_1", torch.cat([enc3, F.max_unpool2d(dec4, i3, 2, 2)], 1))
    dec3 = F.conv2d(dec3, sd[P + "conv_head3_1.weight"])
    dec2 = stage("conv2_1", "SPADE2_1", torch.cat([enc2, F.max_unpool2d(dec3, i2, 2, 2)], 1))
    dec1 = stage("conv1_1", "SPADE1_1", torch.cat([enc1, _up(dec2)], 1))
    dec0 = stage("conv0_1", "SPADE0_1", torch.cat([enc0, _up(dec1)], 1))
    return F.conv2d(dec0, sd[P + "final.weight"], sd[P + "final.bias"])


def spectral_weight(w_orig, u, v, training=True, n_power_iterations=1, eps=1e-12):
    """spectral_norm.py:38-88 (compute_weight): returns (W/sigma, u', v', sigma)."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            for _ in range(n_power_iterations):
                v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
                u = F.normalize(torch.mv(wm, v), dim=0, eps=eps)
    sigma = torch.dot(u, torch.mv(wm, v))
    return w_orig / sigma, u, v, sigma


def discriminator(sd, x, training=True, sync_stats=None, n_blocks=8, spectral=False):
    """models_seg_gan.py:287-300.  ``spectral=True``: every conv / linear weight is
    reparametrised by spectral_norm (keys <name>_orig, <name>_u, <name>_v as written by
    spectral_norm.py:109-142); u/v buffers in ``sd`` are updated in place in training mode."""

    def weight(name):
        if not spectral:
            return sd[name]
        w, u, v, _ = spectral_weight(sd[name + "_orig"], sd[name + "_u"], sd[name + "_v"], training)
        if training:
            sd[name + "_u"], sd[name + "_v"] = u, v
        return w

    out = x
    for i in range(n_blocks):
        p = "conv_blocks.%d.conv_block" % i
        stride = 1 if i % 2 == 0 else 2
        out = F.conv2d(out, weight(p + ".0.weight"), sd[p + ".0.bias"], stride, 1)
        if i != 0:
            out = batch_norm(sd, p + ".1", out, training, sync_stats=sync_stats)
        out = F.leaky_relu(out, 0.2)
    out = F.adaptive_avg_pool2d(out, (6, 6))
    out = F.linear(out.reshape(x.shape[0], -1), weight("fc1.weight"), sd["fc1.bias"])
    out = F.leaky_relu(out, 0.2)
    return F.linear(out, weight("fc2.weight"), sd["fc2.bias"])


# --------------------------------------------------------------------------------------
# losses, metrics, optimiser
# --------------------------------------------------------------------------------------
def stable_bce(x, t):
    """losses.py:133-136."""
    return (x.clamp(min=0) - x * t + (1 + (-x.abs()).exp()).log()).mean()


def bce_dice_loss(x, t):
    """losses.py:280-302 including the NaN/Inf fallback to 2*dice."""
    bce = stable_bce(x, t)
    smooth = 1e-5
    n = t.shape[0]
    p = torch.sigmoid(x).reshape(n, -1)
    tt = t.reshape(n, -1)
    dice = (2.0 * (p * tt).sum(1) + smooth) / (p.sum(1) + tt.sum(1) + smooth)
    dice = 1 - dice.sum() / n
    if torch.isinf(bce) or torch.isnan(bce):
        return 2.0 * dice
    return 0.5 * bce + dice


def iou_score_from_probs(prob: np.ndarray, target: np.ndarray):
    """metrics.py:13-22 on already-computed probabilities (numpy)."""
    smooth = 1e-5
    o = prob > 0.5
    o[np.isnan(prob)] = False
    t = target > 0.5
    return ((o & t).sum() + smooth) / ((o | t).sum() + smooth)


def dice_coef_from_probs(prob: np.ndarray, target: np.ndarray):
    """metrics.py:31-35 on already-computed probabilities (numpy float32 pairwise sums)."""
    smooth = 1e-5
    prob = prob.reshape(-1)
    target = target.reshape(-1)
    inter = (prob * target).sum()
    return (2.0 * inter + smooth) / (prob.sum() + target.sum() + smooth)


def iou_score(logits, target):
    """metrics.py:6-22."""
    return iou_score_from_probs(torch.sigmoid(logits).detach().cpu().numpy(), target.detach().cpu().numpy())


def dice_coef(logits, target):
    """metrics.py:25-35."""
    return dice_coef_from_probs(torch.sigmoid(logits).reshape(-1).detach().cpu().numpy(),
                                target.reshape(-1).detach().cpu().numpy())


def numpy_pairwise_sum_f32(a: np.ndarray) -> np.float32:
    """Pure-Python restatement of NumPy's float32 pairwise summation (used by ndarray.sum on a
    contiguous 1-D float32 array; numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE = 128).
    Small cases only; the CUDA metric kernel reproduces exactly this tree."""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)

    def rec(lo, n):
        if n < 8:
            r = np.float32(0.0) if n == 0 else None
            # numpy: res = 0.; for i: res += a[i]  (starts from 0., exact for the first add)
            r = np.float32(0.0)
            for i in range(n):
                r = np.float32(r + a[lo + i])
            return r
        if n <= 128:
            r = [a[lo + j] for j in range(8)]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] = np.float32(r[j] + a[lo + i + j])
                i += 8
            res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                             np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
            while i < n:
                res = np.float32(res + a[lo + i])
                i += 1
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return np.float32(rec(lo, n2) + rec(lo + n2, n - n2))

    return rec(0, a.size)


class AdamState:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay) restated
    (train_seg_gan.py:452,468; train.py:290 passes config['weight_decay']) preceded by clip_gradient's element clamp
    (srgan_utils.py:186-195).  A parameter whose gradient is None is skipped entirely, as torch does."""

    def __init__(self, names, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.names = list(names)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.t = 0
        self.m = {}
        self.v = {}

    def step(self, sd, grads, grad_clip=None):
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        with torch.no_grad():
            for k in self.names:
                g = grads.get(k)
                if g is None:
                    continue
                if grad_clip is not None:
                    g = g.clamp(-grad_clip, grad_clip)
                if self.weight_decay:
                    g = g.add(sd[k], alpha=self.weight_decay)
                if k not in self.m:
                    self.m[k] = torch.zeros_like(g)
                    self.v[k] = torch.zeros_like(g)
                self.m[k].mul_(b1).add_(g, alpha=1 - b1)
                self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
                sd[k].addcdiv_(self.m[k], denom, value=-(self.lr / bc1))


def trainable_keys(sd):
    return [k for k, v in sd.items() if v.is_floating_point()
            and not k.endswith(("running_mean", "running_var", "_u", "_v"))]


def _leafify(sd):
    for k in trainable_keys(sd):
        sd[k] = sd[k].detach().clone().requires_grad_(True)


def gan_train_step(sd_g, sd_d, opt_g, opt_d, x, target, num_classes=3, sync_stats=None,
                   spectral=False, alpa=1e-4, beta=1e-3, grad_clip=0.8):
    """One iteration of the loop body train_seg_gan.py:188-233 (without .cuda()).

    Mutates sd_g / sd_d / optimiser states in place; returns a dict of the scalars the loop
    produces plus the (post-NaN-scrub) generator logits."""
    _leafify(sd_g)
    _leafify(sd_d)
    gen = unet_r_ss_v2(sd_g, x, True, sync_stats)                         # :188
    gen = torch.where(torch.isnan(gen), torch.zeros_like(gen), gen)       # :190
    out_m = gen[:, 1:num_classes].detach().clone()                        # :191
    tar_m = target[:, 1:num_classes].clone()                              # :192
    loss = bce_dice_loss(gen, target)                                     # :194
    content = F.mse_loss(gen, target)                                     # :195
    iou = iou_score(out_m, tar_m)                                         # :197
    dice = dice_coef(out_m, tar_m)                                        # :198
    seg_d = discriminator(sd_d, gen, True, sync_stats, spectral=spectral)  # :202
    adv_g = F.binary_cross_entropy_with_logits(seg_d, torch.ones_like(seg_d))   # :204
    perceptual = loss + alpa * content + beta * adv_g                     # :205
    gk = trainable_keys(sd_g)
    grads = torch.autograd.grad(perceptual, [sd_g[k] for k in gk], allow_unused=True)   # :207-208
    opt_g.step(sd_g, dict(zip(gk, grads)), grad_clip)                     # :211-215
    hr_d = discriminator(sd_d, target, True, sync_stats, spectral=spectral)          # :217
    sr_d = discriminator(sd_d, gen.detach(), True, sync_stats, spectral=spectral)    # :218
    adv_d = (F.binary_cross_entropy_with_logits(sr_d, torch.zeros_like(sr_d)) +
             F.binary_cross_entropy_with_logits(hr_d, torch.ones_like(hr_d)))        # :221-222
    dk = trainable_keys(sd_d)
    dgrads = torch.autograd.grad(adv_d, [sd_d[k] for k in dk], allow_unused=True)    # :225-226
    opt_d.step(sd_d, dict(zip(dk, dgrads)), grad_clip)                    # :229-233
    return {"loss": float(loss), "content": float(content), "adv_g": float(adv_g),
            "adv_d": float(adv_d), "iou": iou, "dice": dice, "logits": gen.detach(),
            "g_grads": dict(zip(gk, grads)), "d_grads": dict(zip(dk, dgrads))}


def generator_fwd_bwd(sd_g, x, target, sync_stats=None):
    """BASELINE config 1: G forward + BCEDiceLoss + backward (train.py:85-108 inner step)."""
    _leafify(sd_g)
    out = unet_r_ss_v2(sd_g, x, True, sync_stats)
    loss = bce_dice_loss(out, target)
    gk = trainable_keys(sd_g)
    grads = torch.autograd.grad(loss, [sd_g[k] for k in gk], allow_unused=True)
    return out.detach(), loss.detach(), dict(zip(gk, grads))


# --------------------------------------------------------------------------------------
# xResidualBlock (xresidualblock.py:9-33)
# --------------------------------------------------------------------------------------
def xresidual_block_spec(cin=64, planes=64, k=3, sk=9):
    s = [("md.features.0.weight", (planes, cin, k, k)), ("md.features.0.bias", (planes,))]
    s += _bn_keys("md.module.0", planes)
    s += [("md.module.2.weight", (planes, 1, sk, sk)), ("md.module.2.bias", (planes,))]
    s += _bn_keys("md.module.3", planes)
    s += [("conv2.weight", (planes, planes, k, k)), ("conv2.bias", (planes,))]
    s += _bn_keys("bn1", planes)
    return s


def xresidual_block(sd, x, training=True, stride=1):
    """x1 = conv(x); x2 = exp(-(BN(dw9x9(relu(BN(x1)))))^2); y = x1*x2; BN(conv2(y)) + x."""
    k = sd["md.features.0.weight"].shape[-1]
    x1 = F.conv2d(x, sd["md.features.0.weight"], sd["md.features.0.bias"], 1, (k - 1) // 2)
    t = F.relu(batch_norm(sd, "md.module.0", x1, training))
    sk = sd["md.module.2.weight"].shape[-1]
    t = F.conv2d(t, sd["md.module.2.weight"], sd["md.module.2.bias"], 1, (sk - 1) // 2, 1, t.shape[1])
    t = batch_norm(sd, "md.module.3", t, training)
    y = x1 * torch.exp(-(t * t))
    y = F.conv2d(y, sd["conv2.weight"], sd["conv2.bias"], stride, 1)
    return batch_norm(sd, "bn1", y, training) + x


# --------------------------------------------------------------------------------------
# EfficientNet encoder (efficientnet_pytorch/model.py, utils.py) and AttentiveCNN (archs.py:409-466)
# --------------------------------------------------------------------------------------
_EFF_COEFFS = {  # utils.py:153-169: width, depth, resolution
    "efficientnet-b0": (1.0, 1.0, 224), "efficientnet-b1": (1.0, 1.1, 240), "efficientnet-b2": (1.1, 1.2, 260),
    "efficientnet-b3": (1.2, 1.4, 300), "efficientnet-b4": (1.4, 1.8, 380), "efficientnet-b5": (1.6, 2.2, 456),
}
# utils.py:258-263 as (kernel, repeats, in, out, expand, stride); every stage has se_ratio 0.25 and id_skip
_EFF_STAGES = [(3, 1, 32, 16, 1, 1), (3, 2, 16, 24, 6, 2), (5, 2, 24, 40, 6, 2), (3, 3, 40, 80, 6, 2),
               (5, 3, 80, 112, 6, 1), (5, 4, 112, 192, 6, 2), (3, 1, 192, 320, 6, 1)]
EFF_BN_EPS, EFF_BN_MOMENTUM = 1e-3, 0.01      # utils.py:268-269 (momentum = 1 - 0.99, model.py:31)


def _eff_round_filters(f, width, divisor=8):
    """utils.py:56-68."""
    f = f * width
    nf = max(divisor, int(f + divisor / 2) // divisor * divisor)
    if nf < 0.9 * f:
        nf += divisor
    return int(nf)


def efficientnet_blocks(model_name):
    """Per-block dicts in network order (model.py:167-182) plus (stem_out, head_out, image_size)."""
    width, depth, res = _EFF_COEFFS[model_name]
    blocks = []
    for k, r, i, o, e, s in _EFF_STAGES:
        i, o = _eff_round_filters(i, width), _eff_round_filters(o, width)
        r = int(math.ceil(depth * r))
        for j in range(r):
            blocks.append(dict(k=k, cin=i if j == 0 else o, cout=o, expand=e, stride=s if j == 0 else 1,
                               sq=max(1, int((i if j == 0 else o) * 0.25))))
    return blocks, _eff_round_filters(32, width), _eff_round_filters(1280, width), res


def mbconv_spec(p, b):
    """Keys of one MBConvBlock in registration order (model.py:44-62)."""
    mid = b["cin"] * b["expand"]
    k = []
    if b["expand"] != 1:
        k += [(p + "._expand_conv.weight", (mid, b["cin"], 1, 1))] + _bn_keys(p + "._bn0", mid)
    k += [(p + "._depthwise_conv.weight", (mid, 1, b["k"], b["k"]))] + _bn_keys(p + "._bn1", mid)
    k += [(p + "._se_reduce.weight", (b["sq"], mid, 1, 1)), (p + "._se_reduce.bias", (b["sq"],)),
          (p + "._se_expand.weight", (mid, b["sq"], 1, 1)), (p + "._se_expand.bias", (mid,))]
    k += [(p + "._project_conv.weight", (b["cout"], mid, 1, 1))] + _bn_keys(p + "._bn2", b["cout"])
    return k


def efficientnet_spec(model_name, prefix="", num_classes=1000):
    blocks, stem, head, _ = efficientnet_blocks(model_name)
    k = [(prefix + "_conv_stem.weight", (stem, 3, 3, 3))] + _bn_keys(prefix + "_bn0", stem)
    for i, b in enumerate(blocks):
        k += mbconv_spec(prefix + "_blocks.%d" % i, b)
    k += [(prefix + "_conv_head.weight", (head, blocks[-1]["cout"], 1, 1))] + _bn_keys(prefix + "_bn1", head)
    k += [(prefix + "_fc.weight", (num_classes, head)), (prefix + "_fc.bias", (num_classes,))]
    return k


def _static_same_conv(x, w, b, stride, image_size, groups=1):
    """Conv2dStaticSamePadding (utils.py:127-146): the pad comes from the nominal image size, not from x."""
    kk = w.shape[-1]
    o = math.ceil(image_size / stride)
    pad = max((o - 1) * stride + (kk - 1) + 1 - image_size, 0)
    if pad > 0:
        x = F.pad(x, [pad // 2, pad - pad // 2, pad // 2, pad - pad // 2])
    return F.conv2d(x, w, b, stride, 0, 1, groups)


def _swish(x):
    return x * torch.sigmoid(x)


def mbconv_block(sd, p, x, b, image_size, training=True):
    """model.py:64-94 without drop_connect (rate 0 / eval)."""
    def bn(name, t):
        return batch_norm(sd, p + name, t, training, EFF_BN_EPS, EFF_BN_MOMENTUM)

    inp = x
    if b["expand"] != 1:
        x = _swish(bn("._bn0", _static_same_conv(x, sd[p + "._expand_conv.weight"], None, 1, image_size)))
    x = _swish(bn("._bn1", _static_same_conv(x, sd[p + "._depthwise_conv.weight"], None, b["stride"], image_size, x.shape[1])))
    sq = F.adaptive_avg_pool2d(x, 1)
    sq = F.conv2d(_swish(F.conv2d(sq, sd[p + "._se_reduce.weight"], sd[p + "._se_reduce.bias"])),
                  sd[p + "._se_expand.weight"], sd[p + "._se_expand.bias"])
    x = torch.sigmoid(sq) * x
    x = bn("._bn2", F.conv2d(x, sd[p + "._project_conv.weight"]))
    if b["stride"] == 1 and b["cin"] == b["cout"]:
        x = x + inp
    return x


def efficientnet_features(sd, x, model_name, training=True, prefix=""):
    """EfficientNet.extract_features (model.py:202-218), drop_connect off."""
    blocks, _, _, res = efficientnet_blocks(model_name)
    x = _swish(batch_norm(sd, prefix + "_bn0", _static_same_conv(x, sd[prefix + "_conv_stem.weight"], None, 2, res), training,
                          EFF_BN_EPS, EFF_BN_MOMENTUM))
    for i, b in enumerate(blocks):
        x = mbconv_block(sd, prefix + "_blocks.%d" % i, x, b, res, training)
    x = F.conv2d(x, sd[prefix + "_conv_head.weight"])
    return _swish(batch_norm(sd, prefix + "_bn1", x, training, EFF_BN_EPS, EFF_BN_MOMENTUM))


def attentive_cnn_spec(model_name="efficientnet-b2", f_channel=1408):
    return efficientnet_spec(model_name, prefix="eff_conv.") + [("conv_a.weight", (1024, f_channel, 1, 1))]


def attentive_cnn(sd, images, model_name="efficientnet-b2", training=True):
    """archs.py:454-466."""
    res = _EFF_COEFFS[model_name][2]
    r = F.interpolate(images, size=(res, res), mode="bilinear")
    return F.conv2d(efficientnet_features(sd, r, model_name, training, prefix="eff_conv."), sd["conv_a.weight"])


# --------------------------------------------------------------------------------------
# Tiled inference merge (aerial_image_segmentation_api.py:30-217), numpy like the reference
# --------------------------------------------------------------------------------------
def tile_windows(img_h, img_w, p_size, overlap):
    """(h1, w1) per patch in the order patch_gen / patch_merge visit them (:45-126, :141-205)."""
    step = int(math.ceil((1 - overlap) * p_size))
    i_w = int(math.floor((img_w - p_size) / step)) + 1
    i_h = int(math.floor((img_h - p_size) / step)) + 1
    grid = [(i, j) for i in range(i_w) for j in range(i_h)]
    wins = [(j * step, i * step) for i, j in grid]
    wins += [(img_h - j * step - p_size, img_w - i * step - p_size) for i, j in grid]
    wins += [(img_h - j * step - p_size, i * step) for i, j in grid]
    wins += [(j * step, img_w - i * step - p_size) for i, j in grid]
    return wins


def tile_test_probs(base):
    """Deterministic probability maps for the merge fixtures from a coarse [P, C, s, s] float32 seed: x8 block upsampling
    plus an integer ripple (exact float32 arithmetic only, so generator and tests rebuild identical bits)."""
    base = np.asarray(base, dtype=np.float32)
    P, C, s, _ = base.shape
    up = np.repeat(np.repeat(base, 8, axis=2), 8, axis=3)
    p = np.arange(P, dtype=np.int64).reshape(P, 1, 1, 1)
    c = np.arange(C, dtype=np.int64).reshape(1, C, 1, 1)
    y = np.arange(8 * s, dtype=np.int64).reshape(1, 1, -1, 1)
    x = np.arange(8 * s, dtype=np.int64).reshape(1, 1, 1, -1)
    ripple = ((y * 7 + x * 13 + p * 3 + c * 5) % 32).astype(np.float32) / np.float32(128)
    return np.clip(up * np.float32(0.75) + ripple, np.float32(0), np.float32(1)).astype(np.float32)


def _threshold127(m):
    """post_process_resized_mask (:30-42)."""
    m = m.copy()
    m[(m > 127) & (m < 255)] = 255
    m[(m > 0) & (m <= 127)] = 0
    return m


def tile_merge(img_h, img_w, masks, p_size, num_classes, overlap):
    """patch_merge (:129-217) for maps whose size equals the patch size (cv2.resize is then the identity)."""
    wins = tile_windows(img_h, img_w, p_size, overlap)
    out = []
    for c in range(num_classes):
        merged = np.zeros((img_h, img_w))
        div = np.zeros((img_h, img_w))
        for (h1, w1), m in zip(wins, masks):
            u8 = (np.asarray(m[c]) * 255).astype("uint8")
            merged[h1:h1 + p_size, w1:w1 + p_size] += _threshold127(u8) / 255.0
            div[h1:h1 + p_size, w1:w1 + p_size] += 1.0
        div[div == 0] = 1.0
        out.append(_threshold127((np.divide(merged, div) * 255).astype("uint8")))
    return out


# --------------------------------------------------------------------------------------
# bf16-storage emulation of the generator forward (what an ideal bf16 implementation computes)
# --------------------------------------------------------------------------------------
def _q(t):
    return t.bfloat16().float()


def unet_r_ss_v2_bf16_emulated(sd, x, prefix=""):
    """archs.py:623-671 with every stored activation and every conv weight rounded to bf16 (fp32
    accumulation, fp32 BN statistics).  Used to separate "bf16 storage noise amplified by the
    network" from kernel errors: the bf16 CUDA path is compared against this."""
    P = prefix

    def conv(t, w, b=None, pad=1):
        return _q(F.conv2d(t, _q(w), b, 1, pad))

    def bn(p, t, res=None):
        y = F.batch_norm(t, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, 1e-5)
        if res is not None:
            y = y + res
        return _q(F.relu(y))

    def block(p, t):
        r1 = bn(p + ".bn1", conv(t, sd[p + ".conv1.weight"]))
        c2 = conv(r1, sd[p + ".conv2.weight"])
        return bn(p + ".bn2", c2, conv(t, sd[p + ".shortcut.0.weight"], None, 0))

    def spd(p, t):
        seg = conv(t, sd[p + ".x2map.weight"], sd[p + ".x2map.bias"])
        a = _q(F.relu(F.conv2d(seg, _q(sd[p + ".mlp_shared.0.weight"]), sd[p + ".mlp_shared.0.bias"], 1, 1)))
        g = conv(a, sd[p + ".mlp_gamma.weight"], sd[p + ".mlp_gamma.bias"])
        b = conv(a, sd[p + ".mlp_beta.weight"], sd[p + ".mlp_beta.bias"])
        return _q(t * (1 + g) + b)

    def stage(c, s, t):
        return spd(P + s, block(P + c, t))

    def up(t):
        return _q(_up(t))

    e0 = stage("conv0_0", "SPADE0_0", _q(x)); p0, _ = F.max_pool2d(e0, 2, 2, return_indices=True)
    e1 = stage("conv1_0", "SPADE1_0", p0); p1, _ = F.max_pool2d(e1, 2, 2, return_indices=True)
    e2 = stage("conv2_0", "SPADE2_0", p1); p2, i2 = F.max_pool2d(e2, 2, 2, return_indices=True)
    e3 = stage("conv3_0", "SPADE3_0", p2); p3, i3 = F.max_pool2d(e3, 2, 2, return_indices=True)
    e4 = stage("conv4_0", "SPADE4_0", p3); p4, i4 = F.max_pool2d(e4, 2, 2, return_indices=True)
    e5 = conv(stage("conv5_0", "SPADE5_0", p4), sd[P + "conv_head5_0.weight"], None, 0)
    d4 = conv(stage("conv4_1", "SPADE4_1", torch.cat([e4, F.max_unpool2d(e5, i4, 2, 2)], 1)), sd[P + "conv_head4_1.weight"], None, 0)
    d3 = conv(stage("conv3_1", "SPADE3_1", torch.cat([e3, F.max_unpool2d(d4, i3, 2, 2)], 1)), sd[P + "conv_head3_1.weight"], None, 0)
    d2 = stage("conv2_1", "SPADE2_1", torch.cat([e2, F.max_unpool2d(d3, i2, 2, 2)], 1))
    d1 = stage("conv1_1", "SPADE1_1", torch.cat([e1, up(d2)], 1))
    d0 = stage("conv0_1", "SPADE0_1", torch.cat([e0, up(d1)], 1))
    return conv(d0, sd[P + "final.weight"], sd[P + "final.bias"], 0)
