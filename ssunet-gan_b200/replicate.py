"""Data-parallel wrapper with replication callbacks (reference: replicate.py:27-94).

The reference is single-process `nn.DataParallel`: every forward scatters the batch, re-broadcasts
all parameters to every GPU, runs one Python thread per replica and gathers outputs on GPU 0.
Here data parallelism is one process per GPU (launched by torchrun): each rank owns a persistent
replica and its shard of the batch, so there is no per-forward parameter broadcast, scatter or
gather.  `DataParallelWithCallback` keeps the reference's name, `.module` attribute and
`__data_parallel_replicate__(ctx, copy_id)` protocol (copy_id == rank), broadcasts rank 0's
parameters/buffers once at construction, and averages gradients over ranks with one flat NCCL
all-reduce when backward finishes.
"""
import functools

import torch
import torch.distributed as dist
from torch import nn

__all__ = ["CallbackContext", "execute_replication_callbacks", "DataParallelWithCallback", "patch_replication_callback"]


class CallbackContext(object):
    pass


def execute_replication_callbacks(modules, copy_ids=None):
    """Calls `__data_parallel_replicate__(ctx, copy_id)` on every sub-module of every replica; the
    j-th sub-modules of all replicas share one context, the master (copy 0) is visited first."""
    master = modules[0]
    ctxs = [CallbackContext() for _ in master.modules()]
    for i, replica in enumerate(modules):
        cid = i if copy_ids is None else copy_ids[i]
        for j, m in enumerate(replica.modules()):
            if hasattr(m, "__data_parallel_replicate__"):
                m.__data_parallel_replicate__(ctxs[j], cid)


def _dist_ready():
    return dist.is_available() and dist.is_initialized()


class DataParallelWithCallback(nn.Module):
    """`DataParallelWithCallback(module, device_ids=None)`; forward(*inputs) runs the local replica."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0, process_group=None):
        super().__init__()
        self.module = module
        self.device_ids = device_ids
        self.process_group = process_group
        self.rank = dist.get_rank(process_group) if _dist_ready() else 0
        self.world_size = dist.get_world_size(process_group) if _dist_ready() else 1
        self._grad_hook_armed = False
        # True: the arena all-reduce is left RUNNING when backward returns (NCCL's stream); optim.FusedClampAdam.step / zero_grad
        # join it.  train_step.gan_train_step sets it around the generator's backward so that the reduction of G's gradients runs
        # beside the discriminator phase, which does not read them.
        self.async_gradients = False
        self._hook_handles = []
        self._params = [p for p in module.parameters() if p.requires_grad]
        if self.world_size > 1:
            # nn.DataParallel with ONE device calls the module directly (no replicate(), no callbacks): SyncBN then stays on its
            # F.batch_norm path (+eps, num_batches_tracked advances).  Only a real multi-rank wrapper switches it to parallel mode.
            self._broadcast_state()
            execute_replication_callbacks([module], [self.rank])
            for m in module.modules():
                if hasattr(m, "_set_process_group"):
                    m._set_process_group(process_group, self.world_size)
            self._hook_handles = [p.register_hook(self._make_hook()) for p in self._params]

    def _broadcast_state(self):
        with torch.no_grad():
            for t in list(self.module.parameters()) + list(self.module.buffers()):
                dist.broadcast(t.data, 0, group=self.process_group)

    def _remove_hooks(self):
        """Detach this wrapper's gradient hooks (convert_model re-wraps the network: the old wrapper must stop all-reducing)."""
        for h in self._hook_handles:
            h.remove()
        self._hook_handles = []

    def _make_hook(self):
        def hook(grad):
            if not self._grad_hook_armed:
                self._grad_hook_armed = True
                torch.autograd.Variable._execution_engine.queue_callback(self.sync_gradients)
            return grad
        return hook

    def sync_gradients(self):
        """Average .grad over ranks: one all-reduce over a flat fp32 bucket (or in place when the
        parameters already share a flat gradient arena, see optim.FusedClampAdam)."""
        self._grad_hook_armed = False
        if self.world_size <= 1:
            return
        grads = [p.grad for p in self._params if p.grad is not None]
        if not grads:
            return
        from .optim import flat_arena_of
        arena = flat_arena_of(grads)
        if arena is not None:
            # SUM over ranks; the 1/world factor rides into the fused clamp+Adam kernel as its gradient scale (applied before the
            # clamp, as averaging before clip_gradient would) instead of a separate pass over the arena
            if self.async_gradients:
                arena._ssg_pending = dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
            else:
                dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=self.process_group)
            arena._ssg_grad_scale = 1.0 / self.world_size
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.process_group)
        flat.div_(self.world_size)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    def replicate(self, module, device_ids):
        if self.world_size > 1:
            execute_replication_callbacks([module], [self.rank])
        return [module]


def patch_replication_callback(data_parallel):
    """Reference API (replicate.py:70-94): make an existing data-parallel wrapper run the callbacks."""
    assert isinstance(data_parallel, (nn.DataParallel, DataParallelWithCallback))
    if isinstance(data_parallel, DataParallelWithCallback):
        return
    old_replicate = data_parallel.replicate

    @functools.wraps(old_replicate)
    def new_replicate(module, device_ids):
        modules = old_replicate(module, device_ids)
        execute_replication_callbacks(modules)
        return modules

    data_parallel.replicate = new_replicate
