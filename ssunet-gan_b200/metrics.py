"""IoU / Dice metrics (reference: metrics.py:6-35), bit-compatible on identical predicted masks.

iou_score: integer counts on the GPU, float64 ratio on the host (what numpy computes).
dice_coef: the reference sums float32 arrays with NumPy's pairwise tree (blocks of <= 128 summed with
8 strided accumulators, halves combined recursively).  The kernel reproduces the leaf order exactly,
one leaf per thread, and the host combines the leaf sums up the same tree, so the float32 result
is bit-identical to `ndarray.sum()` on the same probabilities."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import call

_leaf_cache = {}


def _leaves(n, device):
    key = (n, str(device))
    hit = _leaf_cache.get(key)
    if hit is None:
        L = _lib.lib()
        cnt = L.ssg_pairwise_leaves_host(n, None, 0)
        host = np.empty(cnt + 1, dtype=np.int64)
        L.ssg_pairwise_leaves_host(n, host.ctypes.data_as(ctypes.c_void_p), cnt + 1)
        hit = (cnt, torch.from_numpy(host).to(device))
        _leaf_cache[key] = hit
    return hit


def _prep(output, target):
    if not torch.is_tensor(output) or not output.is_cuda:
        raise _lib.SsgError("metrics: CUDA tensors required (no CPU path)")
    output = output.detach().contiguous().float()
    target = torch.as_tensor(target, device=output.device).detach().contiguous().float()
    return output, target


# Each metric is a device part (one kernel, no host sync: it can sit inside a captured CUDA graph) and a host part (the float64
# ratio / the top of NumPy's pairwise tree) that runs after the device -> host read of its small result.
def iou_counts(output, target):
    """int64 [intersection, union] of (sigmoid(output) > 0.5, target > 0.5) on the device."""
    output, target = _prep(output, target)
    counts = torch.empty(2, dtype=torch.int64, device=output.device)
    call("ssg_iou_counts", output, target, output.numel(), counts)
    return counts


def iou_from_counts(counts):
    smooth = 1e-5
    inter, union = (counts.cpu() if torch.is_tensor(counts) else counts).numpy()
    return (inter + smooth) / (union + smooth)


def iou_score(output, target):
    return iou_from_counts(iou_counts(output, target))


def dice_leaf_sums(output, target, return_probs=False):
    """Device part of dice_coef: float32 [3, leaves] leaf sums of (p*t, p, t) in NumPy's pairwise order; returns (leaf, n[, probs])."""
    output, target = _prep(output, target)
    n = output.numel()
    cnt, offs = _leaves(n, output.device)
    leaf = torch.empty((3, cnt), dtype=torch.float32, device=output.device)
    probs = torch.empty(n, dtype=torch.float32, device=output.device) if return_probs else None
    call("ssg_dice_leaf_sums", output, target, offs, cnt, leaf, probs)
    return (leaf, n, probs) if return_probs else (leaf, n)


def dice_from_leaves(leaf, n):
    """Host part: combine the leaf sums up NumPy's tree (float32) and form the coefficient."""
    smooth = 1e-5
    host = (leaf.cpu() if torch.is_tensor(leaf) else leaf).numpy()
    L = _lib.lib()
    inter, so, st = [np.float32(L.ssg_pairwise_combine_host(np.ascontiguousarray(host[k]).ctypes.data_as(ctypes.c_void_p), n))
                     for k in range(3)]
    return (2. * inter + smooth) / (so + st + smooth)


def dice_sums(output, target, return_probs=False):
    """float32 (sum(p*t), sum(p), sum(t)) exactly as numpy would compute them on sigmoid(output)."""
    res = dice_leaf_sums(output, target, return_probs)
    leaf, n = res[0], res[1]
    host = leaf.cpu().numpy()
    L = _lib.lib()
    sums = [np.float32(L.ssg_pairwise_combine_host(host[k].ctypes.data_as(ctypes.c_void_p), n)) for k in range(3)]
    if return_probs:
        return sums, res[2]
    return sums


def dice_coef(output, target):
    leaf, n = dice_leaf_sums(output, target)
    return dice_from_leaves(leaf, n)
