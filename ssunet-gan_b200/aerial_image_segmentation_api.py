"""Tiled inference over large rasters (reference: aerial_image_segmentation_api.py:30-217, 376-404).

The reference cuts the raster into overlapping square patches in four sweeps (from the top-left corner, from the
bottom-right corner and the two mixed corners, `patch_gen`), runs the model on ONE patch at a time, pulls every
sigmoid map to the host, and merges them in numpy with a per-class Python loop (`patch_merge`).  Here the patches go
through the model in batches and the merge is two kernels on the device (csrc/tiles.cu): integer vote counters per class
and pixel, then the fp64 divide / scale / 127-threshold of the reference -- the uint8 {0, 255} masks are bit-identical
to `patch_merge` on the same probabilities.

Same function names and return types as the reference.  Supported geometry: model input size == patch size (the
reference additionally lets cv2.resize bridge the two; that path raises here).
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import call


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
        if s != self.p:
            raise _lib.SsgError("patch_merge: model output %d != patch size %d (the cv2.resize bridge of the reference is not "
                                "implemented on the device)" % (s, self.p))
        if self.done + b > len(self.wins):
            raise _lib.SsgError("patch_merge: more patches than windows (%d)" % len(self.wins))
        values = values.contiguous().float()
        call("ssg_mask_vote", values, self.win_dev[self.done:self.done + b], b, c, s, self.h, self.w, int(apply_sigmoid), self.pos, self.cnt)
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


def segmentation_inference(model, img_input, img_patch_set, mask_patch_set, config, gt_mask_flag, batch_size=16):
    """Batched forward of every patch + on-device merge (reference :376-404 runs batch 1 and merges on the host).
    img_patch_set: [P, Cin, S, S] float32 (numpy or tensor), already normalised as `get_patched_input` does.
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
