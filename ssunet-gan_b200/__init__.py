"""ssunet-gan_b200: B200-native (sm_100a) implementation of ssUnet-GAN's data-parallel seg-GAN
training step behind the reference's Python module API.

Host code is PyTorch (device memory, streams, autograd plumbing, torch.distributed); every device
operation on the path is hand-written CUDA in libssunet_b200.so, reached through the C ABI in
include/ssunet_b200.h.  There is no CPU path and no cuDNN/ATen fallback.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
from .ops import set_compute_dtype, compute_dtype, set_conv_impl  # noqa: F401

__all__ = ["ops", "archs", "models_seg_gan", "normalization", "batchnorm", "comm", "replicate", "spectral_norm",
           "losses", "metrics", "srgan_utils", "optim", "train_step", "dataset", "aerial_image_segmentation_api", "xresidualblock",
           "efficientnet_pytorch", "set_compute_dtype", "compute_dtype", "set_conv_impl"]
