"""Tiled inference over large rasters (reference: aerial_image_segmentation_api.py:30-217, 376-404).

The reference cuts the raster into overlapping square patches in four sweeps (from the top-left corner, from the
bottom-right corner and the two mixed corners, `patch_gen`), runs the model on ONE patch at a time, pulls every
sigmoid map to the host, and merges them in numpy with a per-class Python loop (`patch_merge`).  Here the patches go
through the model in batches and the merge is two kernels on the device (csrc/tiles.cu): integer vote counters per class
and pixel, then the fp64 divide / scale / 127-threshold of the reference -- the uint8 {0, 255} masks are bit-identical
to `patch_merge` on the same probabilities.

Same function names and return types as the reference.  When the network's input size differs from the patch size
(config_v1.json: patch_size 1024, input 512) the reference bridges the two with cv2.resize on uint8 data -- every image
patch is shrunk before normalisation (:361), every uint8 mask patch is upsampled before the 127 threshold (:150-152).
Both directions run on the device with OpenCV's exact fixed-point arithmetic (csrc/cvresize.h): `resize_u8` and the
resizing variant of the vote kernel, which never materialises the upsampled map.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import call


_TABLES = {}


def _linear_table(src, dst, axis, device):
    """Per-output-coordinate taps of cv2.resize(..., INTER_LINEAR) on uint8 along one axis: int32 [dst, 4] =
    {index 0, index 1, coef 0, coef 1} with 11-bit coefficients (OpenCV resize.cpp: fx = float((d + 0.5) * scale - 0.5),
    saturate_cast<short>(coef * 2048)).  axis 'x': the fraction is zeroed where a tap would leave the image; axis 'y': the
    fraction is kept and the two ROW indices are clamped (resizeGeneric_Invoker) -- the two differ in rounding."""
    key = (src, dst, axis, str(device))
    hit = _TABLES.get(key)
    if hit is not None:
        return hit
    scale = 1.0 / (float(dst) / float(src))
    f = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if axis == "x":
        lo, hi = s < 0, s >= src - 1
        f[lo], s[lo] = 0.0, 0
        f[hi], s[hi] = 0.0, src - 1
        i0, i1 = s, np.minimum(s + 1, src - 1)
    else:
        i0, i1 = np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1)
    coef = np.clip(np.rint(np.stack([np.float32(1.0) - f, f], 1) * np.float32(2048.0)), -32768, 32767)
    tab = np.concatenate([np.stack([i0, i1], 1), coef], 1).astype(np.int32)
    hit = torch.from_numpy(np.ascontiguousarray(tab)).to(device)
    _TABLES[key] = hit
    return hit


def resize_u8(batch_u8, oh, ow):
    """cv2.resize(patch, (ow, oh)) (default INTER_LINEAR) of every uint8 HWC raster of a CUDA batch [N, H, W, C], bit-exact."""
    if not (torch.is_tensor(batch_u8) and batch_u8.is_cuda and batch_u8.dtype == torch.uint8 and batch_u8.dim() == 4):
        raise _lib.SsgError("resize_u8: expected a CUDA uint8 [N, H, W, C] batch")
    t = batch_u8.contiguous()
    n, h, w, c = t.shape
    if (h, w) == (oh, ow):
        return t
    out = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=t.device)
    area = (h == 2 * oh and w == 2 * ow)
    xt = None if area else _linear_table(w, ow, "x", t.device)
    yt = None if area else _linear_table(h, oh, "y", t.device)
    call("ssg_resize_u8_linear", t, out, n, h, w, c, oh, ow, xt, yt)
    return out


def post_process_resized_mask(resized_mask):
    """In-place 127 threshold of a uint8 mask: (127, 255) -> 255, (0, 127] -> 0 (reference :30-42)."""
    half_th = 127
    resized_mask[(resized_mask > half_th) & (resized_mask < 255)] = 255
    resized_mask[(resized_mask > 0) & (resized_mask <= half_th)] = 0
    return resized_mask


def patch_windows(img_h, img_w, p_size, overlap=0.5):
    """(top, left) of every patch in the reference's order: sweep from (0, 0); sweep back from (H, W); bottom-anchored rows
    with left-anchored columns; top-anchored rows with right-anchored columns (reference :45-126)."""
    step = int(math.ceil((1 - overlap) * p_size))
    i_w = int(math.floor((img_w - p_size) / step)) + 1
    i_h = int(math.floor((img_h - p_size) / step)) + 1
    wins = []
    for i in range(i_w):
        for j in range(i_h):
            wins.append((j * step, i * step))
    for i in range(i_w):
        for j in range(i_h):
            wins.append((img_h - j * step - p_size, img_w - i * step - p_size))
    for i in range(i_w):
        for j in range(i_h):
            wins.append((img_h - j * step - p_size, i * step))
    for i in range(i_w):
        for j in range(i_h):
            wins.append((j * step, img_w - i * step - p_size))
    return wins


def patch_gen(img, mask, p_size, overlap=0.5):
    """Lists of p_size x p_size views of `img` and `mask` (H x W x C arrays), one pair per window."""
    wins = patch_windows(img.shape[0], img.shape[1], p_size, overlap)
    for h1, w1 in wins:
        if h1 < 0 or w1 < 0 or h1 + p_size > img.shape[0] or w1 + p_size > img.shape[1]:
            print('err')
    image_patch = [img[h1:h1 + p_size, w1:w1 + p_size, :] for h1, w1 in wins]
    mask_patch = [mask[h1:h1 + p_size, w1:w1 + p_size, :] for h1, w1 in wins]
    return image_patch, mask_patch


class _Merger:
    """Device-side accumulator of patch votes for one raster."""

    def __init__(self, img_h, img_w, p_size, num_classes, p_overlap, device):
        self.h, self.w, self.p, self.c = img_h, img_w, p_size, num_classes
        self.wins = patch_windows(img_h, img_w, p_size, p_overlap)
        self.win_dev = torch.tensor(self.wins, dtype=torch.int32, device=device).reshape(-1, 2).contiguous()
        self.pos = torch.zeros((num_classes, img_h, img_w), dtype=torch.int32, device=device)
        self.cnt = torch.zeros((img_h, img_w), dtype=torch.int32, device=device)
        self.done = 0

    def add(self, values, apply_sigmoid):
        """values: fp32 CUDA [B, C, S, S] for the next B windows."""
        b, c, s, s2 = values.shape
        if c != self.c or s != s2:
            raise _lib.SsgError("patch_merge: expected [B, %d, S, S] maps, got %s" % (self.c, tuple(values.shape)))
        if self.done + b > len(self.wins):
            raise _lib.SsgError("patch_merge: more patches than windows (%d)" % len(self.wins))
        values = values.contiguous().float()
        win = self.win_dev[self.done:self.done + b]
        if s == self.p:
            call("ssg_mask_vote", values, win, b, c, s, self.h, self.w, int(apply_sigmoid), self.pos, self.cnt)
        else:       # the reference's cv2.resize(mask_u8, (p_size, p_size)) bridge (:150-152), folded into the vote
            area = (s == 2 * self.p)
            xt = None if area else _linear_table(s, self.p, "x", values.device)
            yt = None if area else _linear_table(s, self.p, "y", values.device)
            call("ssg_mask_vote_resized", values, win, b, c, s, self.p, self.h, self.w, int(apply_sigmoid), xt, yt, self.pos, self.cnt)
        self.done += b

    def finish(self):
        out = torch.empty((self.c, self.h, self.w), dtype=torch.uint8, device=self.pos.device)
        call("ssg_mask_finalize", self.pos, self.cnt, self.c, self.h, self.w, out)
        host = out.cpu().numpy()
        return [host[c] for c in range(self.c)]


def patch_merge(img, masks, p_size, config, p_overlap, device="cuda"):
    """masks: sequence of [num_classes, S, S] probability maps in window order (numpy or CUDA tensors).
    Returns a list of num_classes uint8 H x W masks in {0, 255} (reference :129-217)."""
    m = _Merger(img.shape[0], img.shape[1], p_size, config['num_classes'], p_overlap, torch.device(device))
    chunk = 64
    for i in range(0, len(masks), chunk):
        part = masks[i:i + chunk]
        if torch.is_tensor(part):
            vals = part.to(device=m.pos.device, dtype=torch.float32)
        else:
            vals = torch.stack([torch.as_tensor(np.asarray(a), dtype=torch.float32) for a in part]).to(m.pos.device)
        m.add(vals, apply_sigmoid=False)
    return m.finish()


def get_patched_input(img, config, gt_mask_flag=False, device="cuda", chunk=32):
    """`get_patched_input` of the reference (:336-373) with the per-patch work on the device.  img: path (read with
    cv2.imread as the reference does) or the decoded uint8 H x W x 3 raster.  Every patch_size window is shrunk to
    config['input_w'] with cv2.resize's arithmetic, normalised with albumentations' default Normalize() and divided by
    255 once more (the reference's `img.astype('float32') / 255` after Normalize, :367), transposed to CHW.
    Returns (img_input, img_patch_set: CUDA float32 [P, 3, S, S], mask_patch_set: the raw uint8 patches)."""
    from .dataset import DeviceFeed
    if gt_mask_flag:
        raise _lib.SsgError("get_patched_input: ground-truth label rasters are host-side preprocessing outside this package")
    if isinstance(img, str):
        import cv2
        img = cv2.imread(img)
    p_size, size, overlap = config['patch_size'], config['input_w'], config['patch_overlap']
    image_patch, mask_patch = patch_gen(img, img, p_size, overlap)
    feed = DeviceFeed(device=device)
    outs = []
    for i in range(0, len(image_patch), chunk):
        raw = torch.from_numpy(np.ascontiguousarray(np.stack(image_patch[i:i + chunk]))).to(device, non_blocking=True)
        outs.append(feed.images(resize_u8(raw, size, size), nchw=True, post_div=255.0))
    return img, torch.cat(outs, 0) if len(outs) > 1 else outs[0], np.array(mask_patch)


def segmentation_inference(model, img_input, img_patch_set, mask_patch_set, config, gt_mask_flag, batch_size=16):
    """Batched forward of every patch + on-device merge (reference :376-404 runs batch 1 and merges on the host).
    img_patch_set: [P, Cin, S, S] float32 (numpy or tensor), already normalised as `get_patched_input` does; S may differ
    from config['patch_size'] (the maps are then upsampled with cv2.resize's arithmetic inside the vote kernel).
    Returns (all_class_mask, gt_class_mask): lists of uint8 masks per class."""
    patch_size = config['patch_size']
    p_overlap = config['patch_overlap']
    if gt_mask_flag:
        raise _lib.SsgError("segmentation_inference: ground-truth mask conversion (mask_convert, cv2-based) is host-side "
                            "preprocessing outside this package; pass gt_mask_flag=False")
    dev = next(model.parameters()).device
    inp = torch.as_tensor(img_patch_set)
    m = _Merger(img_input.shape[0], img_input.shape[1], patch_size, config['num_classes'], p_overlap, dev)
    with torch.no_grad():
        for i in range(0, inp.shape[0], batch_size):
            logits = model(inp[i:i + batch_size].to(dev, non_blocking=True).float())
            m.add(logits, apply_sigmoid=True)                 # torch.sigmoid(output) folded into the vote kernel
    all_class_mask = m.finish()
    return all_class_mask, all_class_mask
