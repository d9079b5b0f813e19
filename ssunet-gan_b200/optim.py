"""Flat-arena clamp + Adam: one kernel launch updates every parameter of a network.

Parameters, gradients and both Adam moments live in four flat fp32 buffers; `p.data` / `p.grad`
are views into them, so (a) the gradient all-reduce is ONE NCCL call on the gradient arena,
(b) clip_gradient + Adam.step (reference: srgan_utils.py:186-195 + torch.optim.Adam,
train_seg_gan.py:452,468; 170 + 34 parameter tensors) is ONE launch, (c) zero_grad is one memset,
(d) the packed bf16 operands of every convolution weight are refreshed by ONE launch after the step
(`ops.PackRegistry`).

Construction order: the parameters must already be on their CUDA device (`model.cuda()` first).  The
reference builds its optimiser before `.cuda()` (train_seg_gan.py:452 then :474); doing that here raises,
because the arenas would be host memory handed to device kernels.  A later `module.cuda()` / `.to()` /
`.float()` re-creates `p.data` outside the arena: `step()` and `zero_grad()` detect the drift and re-adopt
the parameters (same device) or raise (other device)."""
import torch

from . import ops
from ._lib import SsgError, call


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
        if not params:
            raise SsgError("FusedClampAdam: no trainable parameters")
        dev = params[0].device
        bad = [tuple(p.shape) for p in params if not p.is_cuda or p.device != dev or p.dtype != torch.float32]
        if bad:
            raise SsgError("FusedClampAdam: every parameter must be an fp32 CUDA tensor on ONE device before the optimiser is built "
                           "(call model.cuda() first; the reference builds Adam before .cuda(), train_seg_gan.py:452,474 -- that order "
                           "is not supported here): %d offending parameter(s), first shape %r" % (len(bad), bad[0]))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._clip = grad_clip
        self._pending_clip = None
        self._step = 0
        n = sum(p.numel() for p in params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.packs = ops.PackRegistry()
        self._params = params
        self._offsets = []
        off = 0
        for p in params:
            self._offsets.append(off)
            off += p.numel()
        self._bind(copy_values=True)
        self._step_dev = None          # device-resident step count + hyper-parameters (capturable mode, see make_capturable)
        self._hp_dev = None
        self._hp_host = None
        ops.bump_weight_epoch()

    # ------------------------------------------------------------------------------------------
    # arena bindings
    # ------------------------------------------------------------------------------------------
    def _bind(self, copy_values):
        """Point every p.data / p.grad at its arena slot.  copy_values: adopt the parameter's current values / gradient first."""
        with torch.no_grad():
            for p, off in zip(self._params, self._offsets):
                k = p.numel()
                slot = self.flat_p[off:off + k].view(p.shape)
                if p.data.data_ptr() != slot.data_ptr():
                    if not p.is_cuda or p.device != self.flat_p.device:
                        raise SsgError("FusedClampAdam: parameter of shape %r moved to %s after the optimiser was built (arena on %s)"
                                       % (tuple(p.shape), p.device, self.flat_p.device))
                    if copy_values:
                        slot.copy_(p.data.to(torch.float32))
                    p.data = slot
                g = self.flat_g[off:off + k].view(p.shape)
                if p.grad is None or p.grad.data_ptr() != g.data_ptr() or getattr(p.grad, "_ssg_arena", None) is not self.flat_g:
                    if copy_values and p.grad is not None and p.grad.data_ptr() != g.data_ptr() and p.grad.device == g.device:
                        g.copy_(p.grad)
                    g._ssg_arena = self.flat_g
                    p.grad = g
                p._ssg_packs = self.packs

    def _check_bindings(self):
        """p.data / p.grad must still alias the arenas (Module._apply -- .cuda(), .to(), .float() -- swaps them out while the
        Python-side tags survive).  Pointer comparison per parameter; re-adopts drifted parameters."""
        base_p, base_g = self.flat_p.data_ptr(), self.flat_g.data_ptr()
        for p, off in zip(self._params, self._offsets):
            g = p.grad
            if p.data_ptr() != base_p + 4 * off or g is None or g.data_ptr() != base_g + 4 * off:
                if torch.cuda.is_current_stream_capturing():
                    raise SsgError("FusedClampAdam: parameter storage changed while a CUDA graph is being captured")
                self._bind(copy_values=True)
                ops.bump_weight_epoch()
                return False
        return True

    # ------------------------------------------------------------------------------------------
    def make_capturable(self):
        """Keep the Adam step count AND the hyper-parameters (lr, betas, eps, clip, gradient scale, weight decay) on the device
        so `step()` can be captured in a CUDA graph and replayed (train_step.GraphedGanStep); `sync_hyperparams()` pushes
        param_group / clip changes made after the capture."""
        if self._step_dev is None:
            dev = self.flat_p.device
            self._step_dev = torch.full((1,), float(self._step), dtype=torch.float32, device=dev)
            self._hp_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        return self

    def _hyper(self, clip, grad_scale):
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        return (float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(clip) if clip is not None else 0.0,
                float(grad_scale), float(group.get("weight_decay", 0.0) or 0.0), 0.0)

    def sync_hyperparams(self, clip="keep", grad_scale="keep"):
        """Write the current param_group values into the device copy when they changed (never during capture)."""
        if self._hp_dev is None:
            return
        old = self._hp_host
        c = (old[4] if old else 0.0) if clip == "keep" else (float(clip) if clip is not None else 0.0)
        s = (old[5] if old else 1.0) if grad_scale == "keep" else float(grad_scale)
        hp = self._hyper(c if c else None, s)
        if hp != old:
            if torch.cuda.is_current_stream_capturing():
                raise SsgError("FusedClampAdam: hyper-parameters changed during CUDA-graph capture")
            self._hp_dev.copy_(torch.tensor(hp, dtype=torch.float32))
            self._hp_host = hp

    def defer_clip(self, grad_clip):
        self._pending_clip = grad_clip

    @torch.no_grad()
    def clamp_weights(self, clip):
        """`for p in model.parameters(): p.data.clamp_(-clip, clip)` (train.py:111-112) as ONE launch on the parameter arena.
        Like the reference's in-place clamp it lands between forward and backward, so the backward pass that follows reads
        the clamped values (packed-weight caches are invalidated)."""
        self._check_bindings()
        call("ssg_clamp_", self.flat_p, self.flat_p.numel(), float(clip))
        ops.bump_weight_epoch()

    def wait_gradients(self):
        """Join a gradient all-reduce that a data-parallel wrapper left running on the arena (replicate.py: `async_gradients`):
        the current stream waits for it (capturable: an event dependency, not a host block)."""
        work = getattr(self.flat_g, "_ssg_pending", None)
        if work is not None:
            work.wait()
            self.flat_g._ssg_pending = None

    def zero_grad(self, set_to_none=False):
        self._check_bindings()
        self.wait_gradients()
        self.flat_g.zero_()

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None):
        self._check_bindings()
        self.wait_gradients()
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        self._step += 1
        clip = self._pending_clip if self._pending_clip is not None else self._clip
        self._pending_clip = None
        if grad_scale is None:
            # a data-parallel wrapper that SUMMED the gradient arena over ranks leaves the 1/world factor here (replicate.py)
            grad_scale = getattr(self.flat_g, "_ssg_grad_scale", 1.0)
        wd = float(group.get("weight_decay", 0.0) or 0.0)
        n = self.flat_p.numel()
        if self._step_dev is not None:
            self.sync_hyperparams(clip, grad_scale)
            call("ssg_clamp_adam_hp_dev", self.flat_p, self.flat_g, self.flat_m, self.flat_v, n, self._hp_dev, self._step_dev)
        else:
            call("ssg_clamp_adam_wd", self.flat_p, self.flat_g, self.flat_m, self.flat_v, n, float(group["lr"]),
                 float(b1), float(b2), float(group["eps"]), 1.0 - b1 ** self._step, 1.0 - b2 ** self._step,
                 float(clip) if clip is not None else 0.0, float(grad_scale), wd)
        ops.bump_weight_epoch(source=self.packs)      # only THIS optimiser's weights changed ...
        self.packs.refresh()                          # ... and all their packed operands are rebuilt here, in one launch

    # ------------------------------------------------------------------------------------------
    # checkpointing: torch.optim.Adam's layout (state[i] = {step, exp_avg, exp_avg_sq} per parameter index)
    # ------------------------------------------------------------------------------------------
    def state_dict(self):
        step = int(round(float(self._step_dev))) if self._step_dev is not None else self._step
        state = {}
        if step > 0:
            for i, (p, off) in enumerate(zip(self._params, self._offsets)):
                k = p.numel()
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": self.flat_m[off:off + k].view(p.shape).clone(),
                            "exp_avg_sq": self.flat_v[off:off + k].view(p.shape).clone()}
        groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": list(range(len(self._params)))} for g in self.param_groups]
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, sd):
        for g, new in zip(self.param_groups, sd["param_groups"]):
            for k, v in new.items():
                if k != "params":
                    g[k] = v
        step = 0
        self.flat_m.zero_()
        self.flat_v.zero_()
        for i, st in sd.get("state", {}).items():
            i = int(i)
            p, off = self._params[i], self._offsets[i]
            k = p.numel()
            self.flat_m[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self.flat_v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(round(float(st["step"]))))
        self._step = step
        if self._step_dev is not None:
            self._step_dev.fill_(float(step))
