"""Data feed for the training / validation steps (reference: dataset.py:10-144 and the albumentations pipelines built at
train_seg_gan.py:363-382) -- SURVEY §8f row 3.

The reference decodes with cv2 in DataLoader workers, runs albumentations on the CPU, converts to float32 CHW and ships
fp32 tensors through `.cuda()`.  Here the host keeps the rasters as the uint8 HWC arrays cv2 returns, they cross PCIe as
bytes (4x less traffic), and ONE kernel per tensor does Normalize (+ the Flip augmentation) and writes either the
reference's NCHW fp32 layout or, for the image, the channel-padded NHWC activation the first convolution reads -- so the
NCHW -> NHWC pass at the network entry disappears as well.

In scope: `Dataset` (file layout and decode order of dataset.py:95-144, minus the transform), Normalize, Flip, the
mask binarisation.  Out of scope (stay on the host if wanted, they commute with this feed): Rotate, HueSaturationValue,
RandomBrightnessContrast, Resize (interpolating augmentations of albumentations, which is not installed here: nothing
to pin them against).
"""
import os

import numpy as np
import torch

from . import ops
from ._lib import SsgError, call, dtype_code

IMAGENET_MEAN = [0.485, 0.456, 0.406]      # train_seg_gan.py:363
IMAGENET_STD = [0.229, 0.224, 0.225]       # train_seg_gan.py:364
FLIP_NONE, FLIP_X, FLIP_Y, FLIP_XY = 0, 1, 2, 3


def flip_code_from_cv2(d):
    """cv2.flip's flipCode (what albumentations' Flip draws from {-1, 0, 1}) -> this module's bit code."""
    return {1: FLIP_X, 0: FLIP_Y, -1: FLIP_XY}[int(d)]


class Dataset(torch.utils.data.Dataset):
    """dataset.py:10-144 with the same constructor; `__getitem__` returns the decoded uint8 rasters
    `(ori_img, img_u8 [H,W,C], mask_u8 [H,W,num_classes], masks, {'img_id': id})` -- the float conversion, transpose and
    Normalize / Flip happen on the device in `DeviceFeed`.  A host `transform` (albumentations-style callable taking
    image=, mask=) is still applied when given, before the feed."""

    def __init__(self, img_ids, img_dir, mask_dir, img_ext, mask_ext, num_classes, input_channels=3, transform=None, from_file=None):
        self.img_ids = img_ids
        self.img_dir = img_dir
        self.mask_dir = mask_dir
        self.img_ext = img_ext
        self.mask_ext = mask_ext
        self.num_classes = num_classes
        self.input_channels = input_channels
        self.transform = transform
        self.from_file = from_file

    def __len__(self):
        return len(self.img_ids)

    def __getitem__(self, idx):
        import cv2
        img_id = self.img_ids[idx]
        if self.input_channels == 3:                                           # dataset.py:98-102
            img = cv2.imread(os.path.join(self.img_dir, img_id + self.img_ext)) if self.from_file is None else self.from_file[img_id]["img"]
        else:                                                                  # :103-106
            img = cv2.imread(os.path.join(self.img_dir, img_id + self.img_ext), cv2.IMREAD_GRAYSCALE)[..., None]
        ori_img = img
        if self.num_classes == 1:                                              # :109-113
            mask = cv2.imread(os.path.join(self.mask_dir, img_id + self.mask_ext), cv2.IMREAD_GRAYSCALE)[..., None]
        else:                                                                  # :126-132: raw PNG bytes; /255 happens in the feed
            mask = np.dstack([cv2.imread(os.path.join(self.mask_dir, str(i), img_id + self.mask_ext), cv2.IMREAD_GRAYSCALE)[..., None]
                              for i in range(self.num_classes)])
        if self.transform is not None:
            aug = self.transform(image=img, mask=mask)
            img, mask = aug["image"], aug["mask"]
        return ori_img, np.ascontiguousarray(img), np.ascontiguousarray(mask), [], {"img_id": img_id}


class DeviceFeed:
    """Normalize(mean, std, max_pixel_value) + Flip + float / layout conversion of a uint8 batch on the GPU.

    images(img_u8 [N,H,W,C]) -> NHWC compute-dtype activation with `ops.thin_pad(C)` stored channels (default, what
    `archs.*` / `Generator` consume without another pass) or the reference's NCHW fp32 tensor (`nchw=True`).
    masks(mask_u8 [N,H,W,K], binarise=True) -> NCHW fp32 target; `binarise` is dataset.py:128-131's
    `(png / 255.0).astype('uint8')` for the multi-class layout (False = the single-class `mask / 1.0` of :112)."""

    def __init__(self, mean=IMAGENET_MEAN, std=IMAGENET_STD, max_pixel_value=255.0, device="cuda"):
        if not torch.cuda.is_available():
            raise SsgError("DeviceFeed: no CUDA device; ssunet-gan_b200 has no CPU path")
        self.device = torch.device(device)
        m = np.array(mean, dtype=np.float32) * np.float32(max_pixel_value)     # albumentations functional.normalize, float32
        s = np.array(std, dtype=np.float32) * np.float32(max_pixel_value)
        self.sub = torch.from_numpy(m).to(self.device)
        self.mul = torch.from_numpy(np.reciprocal(s, dtype=np.float32)).to(self.device)
        self.h2d_bytes = 0

    def _u8(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        if t.dtype != torch.uint8 or t.dim() != 4:
            raise SsgError("DeviceFeed: expected a uint8 [N,H,W,C] batch, got %s %s" % (t.dtype, tuple(t.shape)))
        if not t.is_cuda:
            self.h2d_bytes += t.numel()
            t = t.to(self.device, non_blocking=True)
        return t.contiguous()

    def _flip(self, codes, n):
        if codes is None:
            return None
        c = torch.as_tensor(codes, dtype=torch.int32).reshape(-1)
        if c.numel() != n:
            raise SsgError("DeviceFeed: one flip code per sample expected")
        return c.to(self.device, non_blocking=True)

    def images(self, img_u8, flip_codes=None, nchw=False, post_div=0.0):
        """post_div: divide the normalised values by this (float32) -- `get_patched_input` divides by 255 once more
        (aerial_image_segmentation_api.py:367)."""
        t = self._u8(img_u8)
        n, h, w, c = t.shape
        if c > self.sub.numel():
            raise SsgError("DeviceFeed: %d bands but %d mean/std entries" % (c, self.sub.numel()))
        fc = self._flip(flip_codes, n)
        if nchw:
            out = torch.empty((n, c, h, w), dtype=torch.float32, device=self.device)
            call("ssg_feed_image_u8", t, out, dtype_code(torch.float32), 1, n, h, w, c, c, self.sub, self.mul, float(post_div), fc)
            return out
        dt = ops.compute_dtype()
        cs = ops.thin_pad(c)
        out = ops.empty_nhwc(n, cs, h, w, dt, self.device)
        call("ssg_feed_image_u8", t, out, dtype_code(dt), 0, n, h, w, c, cs, self.sub, self.mul, float(post_div), fc)
        return out

    def masks(self, mask_u8, flip_codes=None, binarise=True):
        t = self._u8(mask_u8)
        n, h, w, k = t.shape
        fc = self._flip(flip_codes, n)
        out = torch.empty((n, k, h, w), dtype=torch.float32, device=self.device)
        if binarise:
            call("ssg_feed_mask_u8", t, out, n, h, w, k, fc)
        else:       # mask.astype('float32') / 1.0 (dataset.py:112,123): the identity normalisation
            one = torch.ones(k, dtype=torch.float32, device=self.device)
            call("ssg_feed_image_u8", t, out, dtype_code(torch.float32), 1, n, h, w, k, k, torch.zeros_like(one), one, 0.0, fc)
        return out

    def __call__(self, img_u8, mask_u8, flip_codes=None, nchw=False):
        return self.images(img_u8, flip_codes, nchw), self.masks(mask_u8, flip_codes)
