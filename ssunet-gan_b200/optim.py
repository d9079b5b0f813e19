"""Flat-arena clamp + Adam: one kernel launch updates every parameter of a network.

Parameters, gradients and both Adam moments live in four flat fp32 buffers; `p.data` / `p.grad`
are views into them, so (a) the gradient all-reduce is ONE NCCL call on the gradient arena,
(b) clip_gradient + Adam.step (reference: srgan_utils.py:186-195 + torch.optim.Adam,
train_seg_gan.py:452,468; 170 + 34 parameter tensors) is ONE launch, (c) zero_grad is one memset."""
import torch

from . import ops
from ._lib import call


def flat_arena_of(grads):
    """If every tensor in `grads` is a view into one registered flat arena, return that arena."""
    arena = getattr(grads[0], "_ssg_arena", None)
    if arena is None:
        return None
    for g in grads:
        if getattr(g, "_ssg_arena", None) is not arena:
            return None
    total = sum(g.numel() for g in grads)
    return arena if total == arena.numel() else None


class FusedClampAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) semantics (L2 weight decay as in
    train.py:290; no amsgrad)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_clip=None, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._clip = grad_clip
        self._pending_clip = None
        self._step = 0
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                g = self.flat_g[off:off + k].view(p.shape)
                g._ssg_arena = self.flat_g
                p.grad = g
                off += k
        self._params = params
        self._step_dev = None          # device-resident step count (capturable mode, see make_capturable)
        ops.bump_weight_epoch()

    def make_capturable(self):
        """Keep the Adam step count on the device so `step()` can be captured in a CUDA graph and replayed
        (train_step.GraphedGanStep)."""
        if self._step_dev is None:
            self._step_dev = torch.full((1,), float(self._step), dtype=torch.float32, device=self.flat_p.device)
        return self

    def defer_clip(self, grad_clip):
        self._pending_clip = grad_clip

    @torch.no_grad()
    def clamp_weights(self, clip):
        """`for p in model.parameters(): p.data.clamp_(-clip, clip)` (train.py:111-112) as ONE launch on the parameter arena.
        Like the reference's in-place clamp it lands between forward and backward, so the backward pass that follows reads
        the clamped values (packed-weight caches are invalidated)."""
        call("ssg_clamp_", self.flat_p, self.flat_p.numel(), float(clip))
        ops.bump_weight_epoch()

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p in self._params:   # re-attach views if autograd replaced them
            if p.grad is None or getattr(p.grad, "_ssg_arena", None) is not self.flat_g:
                self._reattach()
                break

    def _reattach(self):
        off = 0
        for p in self._params:
            k = p.numel()
            g = self.flat_g[off:off + k].view(p.shape)
            g._ssg_arena = self.flat_g
            if p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
            p.grad = g
            off += k

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        for p in self._params:
            if p.grad is None or getattr(p.grad, "_ssg_arena", None) is not self.flat_g:
                self._reattach()
                break
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        self._step += 1
        clip = self._pending_clip if self._pending_clip is not None else self._clip
        self._pending_clip = None
        wd = float(group.get("weight_decay", 0.0) or 0.0)
        if self._step_dev is not None:
            call("ssg_clamp_adam_wd_dev", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_p.numel(), float(group["lr"]),
                 float(b1), float(b2), float(group["eps"]), self._step_dev, float(clip) if clip is not None else 0.0, float(grad_scale),
                 wd)
        else:
            call("ssg_clamp_adam_wd", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_p.numel(), float(group["lr"]),
                 float(b1), float(b2), float(group["eps"]), 1.0 - b1 ** self._step, 1.0 - b2 ** self._step,
                 float(clip) if clip is not None else 0.0, float(grad_scale), wd)
        ops.bump_weight_epoch()
