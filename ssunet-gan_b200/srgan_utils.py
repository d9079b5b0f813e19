"""clip_gradient / AverageMeter (reference: srgan_utils.py:165-195)."""
from ._lib import call


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def clip_gradient(optimizer, grad_clip):
    """Element-wise clamp of every .grad to [-grad_clip, grad_clip], in place (srgan_utils.py:186-195).
    With a FusedClampAdam optimiser the clamp is deferred into its single fused step kernel."""
    if hasattr(optimizer, "defer_clip"):
        optimizer.defer_clip(grad_clip)
        return
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p.grad is not None:
                g = p.grad.data
                if not g.is_contiguous():
                    raise RuntimeError("clip_gradient: non-contiguous gradient")
                call("ssg_clamp_", g, g.numel(), float(grad_clip))
