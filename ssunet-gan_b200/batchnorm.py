"""SynchronizedBatchNorm{1,2,3}d + convert_model (reference: batchnorm.py:40-361).

Per-channel sum / sum-of-squares come from one streaming CUDA kernel; across GPUs they are summed
by ONE NCCL all-reduce of 2*C fp64 values per layer (forward) and one more in backward, replacing
the reference's Python-thread queue rendezvous + ReduceAddCoalesced/Broadcast through GPU 0.
Arithmetic follows _compute_mean_std (batchnorm.py:115-127), including clamp(eps) (not +eps) in
parallel mode and "num_batches_tracked is never incremented on the parallel path"."""
import torch
from torch.nn.modules.batchnorm import _BatchNorm

from . import ops
from ._lib import ACT_NONE
from .comm import SyncMaster
from .nn_layers import BatchNorm2d
from .replicate import DataParallelWithCallback

__all__ = ["SynchronizedBatchNorm1d", "SynchronizedBatchNorm2d", "SynchronizedBatchNorm3d", "convert_model"]


class _SynchronizedBatchNorm(_BatchNorm):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self._sync_master = SyncMaster(self._data_parallel_master)
        self._is_parallel = False
        self._parallel_id = None
        self._slave_pipe = None
        self._group = None
        self._world = 1

    # -- reference hook protocol (batchnorm.py:82-90); copy_id is the rank in this framework
    def __data_parallel_replicate__(self, ctx, copy_id):
        self._is_parallel = True
        self._parallel_id = copy_id
        if copy_id == 0:
            ctx.sync_master = self._sync_master
        elif hasattr(ctx, "sync_master"):
            self._slave_pipe = ctx.sync_master.register_slave(copy_id)

    def _set_process_group(self, group, world):
        # None means "the default group" to torch.distributed, but `ops.batch_norm(group=None)` means "single device":
        # resolve it here so a default-constructed DataParallelWithCallback really synchronises the statistics
        if group is None and world > 1:
            import torch.distributed as dist
            group = dist.group.WORLD
        self._group, self._world = group, world

    def _data_parallel_master(self, intermediates):
        """Thread-level master callback kept for API parity: reduces (sum, ssum, count) messages."""
        total = sum(m[1][2] for m in intermediates)
        s = sum(m[1][0] for m in intermediates)
        ss = sum(m[1][1] for m in intermediates)
        return [(ident, (s, ss, total)) for ident, _ in intermediates]

    def _to4d(self, x):
        return x

    def forward(self, input, residual=None, act=ACT_NONE, slope=0.0, sums=None):
        self._check_input_dim(input)
        shape = input.shape
        x = self._to4d(input)
        parallel = self._is_parallel and self.training
        # F.batch_norm path (batchnorm.py:52-55) advances num_batches_tracked; the parallel path never does
        nbt = self.num_batches_tracked if (self.training and not parallel and self.num_batches_tracked is not None) else None
        y = ops.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, self.training,
                           self.momentum, self.eps, residual, act, slope,
                           self._group if parallel else None, sync_quirk=parallel, sums=sums, nbt=nbt)
        if y.shape != shape:
            y = ops.to_nchw_f32(y).reshape(shape)
        return y


class SynchronizedBatchNorm1d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() not in (2, 3):
            raise ValueError("expected 2D or 3D input (got {}D input)".format(input.dim()))

    def _to4d(self, x):
        return x.reshape(x.shape[0], x.shape[1], -1, 1)


class SynchronizedBatchNorm2d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError("expected 4D input (got {}D input)".format(input.dim()))


class SynchronizedBatchNorm3d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(input.dim()))

    def _to4d(self, x):
        return x.reshape(x.shape[0], x.shape[1], x.shape[2], -1)


def convert_model(module):
    """Replace every BatchNorm{1,2,3}d by its synchronised twin, sharing running stats and cloning the
    affine parameters (batchnorm.py:320-361).  A DataParallel wrapper becomes DataParallelWithCallback."""
    if isinstance(module, DataParallelWithCallback):
        # already this package's wrapper: convert the wrapped network and reuse the wrapper's configuration (a second wrapper
        # around the first would register a second set of gradient hooks: two all-reduces per backward)
        module._remove_hooks()
        return DataParallelWithCallback(convert_model(module.module), device_ids=module.device_ids, process_group=module.process_group)
    if isinstance(module, torch.nn.DataParallel):
        return DataParallelWithCallback(convert_model(module.module), device_ids=module.device_ids)
    mod = module
    for src, dst in ((torch.nn.BatchNorm1d, SynchronizedBatchNorm1d), (torch.nn.BatchNorm2d, SynchronizedBatchNorm2d),
                     (torch.nn.BatchNorm3d, SynchronizedBatchNorm3d)):
        if isinstance(module, src):
            mod = dst(module.num_features, module.eps, module.momentum, module.affine)
            mod.running_mean = module.running_mean
            mod.running_var = module.running_var
            if module.affine:
                mod.weight.data = module.weight.data.clone().detach()
                mod.bias.data = module.bias.data.clone().detach()
            mod.to(module.running_mean.device)
    for name, child in module.named_children():
        mod.add_module(name, convert_model(child))
    return mod
