"""One G+D iteration of the seg-GAN fine-tuning loop (reference: train_seg_gan.py:182-233), one
process per GPU.

`gan_train_step` is a line-by-line mirror of the reference loop body driving this package's modules
(the same call sequence the reference makes, so it also works with `torch.optim.Adam` +
`clip_gradient`).  Differences from the reference that do NOT change results:
  * the discriminator's parameters do not accumulate gradients during the generator step — the
    reference computes them and `optimizer_d.zero_grad()` throws them away (train_seg_gan.py:225);
  * BCEDiceLoss and MSELoss are one fused pass;
  * the IoU/Dice metrics can be skipped (`with_metrics=False`) or left half-done on the device (`with_metrics="device"` +
    `finish_metrics`): their host halves force a sync, which a captured CUDA graph cannot contain.
Data parallelism: wrap G and D in `replicate.DataParallelWithCallback` (after
`batchnorm.convert_model` for SyncBN); gradients are averaged over ranks when backward finishes.
"""
import os
from collections import OrderedDict

import torch

from . import ops
from .losses import BCEDiceAndContentLoss
from . import metrics
from .metrics import dice_coef, iou_score
from .srgan_utils import clip_gradient

ALPA = 1e-4      # train_seg_gan.py:172
BETA = 1e-3      # train_seg_gan.py:173
GRAD_CLIP = 0.8  # train_seg_gan.py:174
OVERLAP_GRAD_SYNC = os.environ.get("SSG_OVERLAP_GRAD_SYNC", "1") != "0"      # see gan_train_step


def _set_requires_grad(module, flag, saved=None):
    if flag is False:
        saved = [(p, p.requires_grad) for p in module.parameters()]
        for p, _ in saved:
            p.requires_grad_(False)
        return saved
    for p, r in saved:
        p.requires_grad_(r)
    return None


def gan_train_step(generator, discriminator, optimizer_g, optimizer_d, input, target, num_classes=3,
                   with_metrics=True, grad_clip=GRAD_CLIP, alpa=ALPA, beta=BETA):
    """input: N x Cin x H x W fp32 CUDA, target: N x num_classes x H x W fp32 {0,1} CUDA.
    Returns an OrderedDict of 0-dim tensors / python scalars: loss, content, adv_g, adv_d, iou, dice."""
    criterion = BCEDiceAndContentLoss()

    # ---- generator update (train_seg_gan.py:188-215) ----
    generator_output = generator(input)
    generator_output = ops.nan_to_zero(generator_output)                       # :190
    loss, content_loss = criterion(generator_output, target)                   # :194-195
    iou = dice = None
    pending = None
    if with_metrics:
        out_m = generator_output[:, 1:num_classes].detach().contiguous()       # :191
        tar_m = target[:, 1:num_classes].contiguous()                          # :192
        if with_metrics == "device":
            # device halves only (counts / pairwise leaf sums): no host sync here, so the step can be captured in a CUDA graph;
            # `finish_metrics` completes them after the step's results are read back
            pending = (metrics.iou_counts(out_m, tar_m),) + metrics.dice_leaf_sums(out_m, tar_m)
        else:
            iou = iou_score(out_m, tar_m)                                      # :197
            dice = dice_coef(out_m, tar_m)                                     # :198
    saved = _set_requires_grad(discriminator, False)
    seg_discriminated = discriminator(generator_output)                        # :202
    _set_requires_grad(discriminator, True, saved)
    adversarial_loss = ops.bce_with_logits_const(seg_discriminated, 1.0)       # :204
    perceptual_loss = loss + alpa * content_loss + beta * adversarial_loss     # :205
    optimizer_g.zero_grad()                                                    # :207
    # Data parallel (replicate.DataParallelWithCallback on a flat-arena optimiser): the all-reduce of G's gradients is left
    # running on NCCL's stream and G's clamp + Adam moves behind the discriminator phase -- nothing in :217-226 reads G's
    # parameters or gradients (generator_output is already computed), so the result is the same and the reduction overlaps
    # D's two forwards and its backward.  OVERLAP_GRAD_SYNC = False keeps the reference's order.
    overlap = (OVERLAP_GRAD_SYNC and getattr(generator, "world_size", 1) > 1 and hasattr(generator, "async_gradients")
               and hasattr(optimizer_g, "wait_gradients"))
    if overlap:
        generator.async_gradients = True
    try:
        perceptual_loss.backward()                                             # :208
    finally:
        if overlap:
            generator.async_gradients = False

    def update_generator():
        if grad_clip is not None:
            clip_gradient(optimizer_g, grad_clip)                              # :211-212
        optimizer_g.step()                                                     # :215

    if not overlap:
        update_generator()
    adv_g = adversarial_loss.detach()

    # ---- discriminator update (train_seg_gan.py:217-233) ----
    hr_discriminated = discriminator(target)                                   # :217
    sr_discriminated = discriminator(generator_output.detach())                # :218
    adversarial_loss = ops.bce_with_logits_const(sr_discriminated, 0.0) + \
        ops.bce_with_logits_const(hr_discriminated, 1.0)                       # :221-222
    optimizer_d.zero_grad()                                                    # :225
    overlap_d = overlap and getattr(discriminator, "world_size", 1) > 1 and hasattr(discriminator, "async_gradients")
    if overlap_d:
        discriminator.async_gradients = True       # D's reduction runs beside G's deferred clamp + Adam + operand re-packing
    try:
        adversarial_loss.backward()                                            # :226
    finally:
        if overlap_d:
            discriminator.async_gradients = False
    if overlap:
        update_generator()
    if grad_clip is not None:
        clip_gradient(optimizer_d, grad_clip)                                  # :229-230
    optimizer_d.step()                                                         # :233
    out = OrderedDict([("loss", loss.detach()), ("content", content_loss.detach()), ("adv_g", adv_g),
                       ("adv_d", adversarial_loss.detach()), ("iou", iou), ("dice", dice),
                       ("logits", generator_output.detach())])
    if pending is not None:
        out["metric_parts"] = pending
    return out


def finish_metrics(out):
    """Host halves of iou_score / dice_coef (metrics.py:6-35) for a step run with with_metrics="device": reads the int64 counts
    and the float32 leaf sums back and fills out["iou"], out["dice"] (bit-identical to the one-call metrics)."""
    counts, leaf, n = out["metric_parts"]
    out["iou"] = metrics.iou_from_counts(counts)
    out["dice"] = metrics.dice_from_leaves(leaf, n)
    return out


def generator_fwd_bwd(generator, input, target):
    """BASELINE config 1: generator forward + BCEDiceLoss + backward (train.py:85-108 without the optimiser)."""
    out = generator(input)
    loss = ops.seg_losses(out, target)[0]
    loss.backward()
    return out.detach(), loss.detach()


class GraphedGanStep:
    """The whole G+D iteration captured once in a CUDA graph and replayed: one launch per step instead of ~2500
    (the eager step is CPU-launch-bound below ~50 ms of device work, e.g. at batch 8 per GPU).

    Static device buffers hold the inputs; `__call__(input, target)` copies the new batch into them (device or pinned-host
    sources) and replays.  The optimisers must be `optim.FusedClampAdam` (flat gradient arenas, device-resident step
    count); BatchNorm running statistics, Adam moments, parameters and the NCCL all-reduces (SyncBN statistics, gradient
    arenas) are all updated by the replayed kernels exactly as in the eager step.  Construction leaves parameters, buffers
    and optimiser state exactly as it found them (the warm-up iterations are rolled back).  Returns the same OrderedDict of 0-dim
    tensors as `gan_train_step` (static outputs: read them before the next call).  The device halves of iou_score / dice_coef
    (train_seg_gan.py:197-198) ARE part of the graph (`with_metrics=True`, the default); `metrics()` reads the counts / leaf sums
    back and finishes them on the host.  Learning rate, betas, eps and clip live in device memory and are refreshed from the
    optimisers' param_groups before every replay, so LR schedulers keep working."""

    def __init__(self, generator, discriminator, optimizer_g, optimizer_d, batch_shape, target_shape=None, num_classes=3,
                 warmup=3, with_metrics=True, **kw):
        from . import _lib
        dev = next(generator.parameters()).device
        self.g, self.d, self.og, self.od = generator, discriminator, optimizer_g, optimizer_d
        self.kw = dict(kw, num_classes=num_classes, with_metrics="device" if with_metrics else False)
        self.x = torch.zeros(batch_shape, dtype=torch.float32, device=dev)
        tshape = target_shape or (batch_shape[0], num_classes, batch_shape[2], batch_shape[3])
        self.t = torch.zeros(tshape, dtype=torch.float32, device=dev)
        self.t[:, 0].fill_(1.0)
        optimizer_g.make_capturable()
        optimizer_d.make_capturable()
        # Building the graph must not train the model: the warm-up iterations below run real optimiser steps on the static
        # (zero) batch, so every piece of state they touch is snapshotted here and restored IN PLACE after the capture
        # (same storage, so the captured kernels keep pointing at it).
        snap = []
        for opt in (optimizer_g, optimizer_d):
            snap += [(t, t.clone()) for t in (opt.flat_p, opt.flat_m, opt.flat_v, opt._step_dev)]
        for mod in (generator, discriminator):
            snap += [(b, b.clone()) for b in mod.buffers()]
            snap += [(q.data, q.data.clone()) for q in mod.parameters() if not q.requires_grad]
        steps0 = (optimizer_g._step, optimizer_d._step)
        # warm up on a side stream (allocator, lazy kernel attributes, NCCL communicators), then capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                gan_train_step(self.g, self.d, self.og, self.od, self.x, self.t, **self.kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # Packed conv operands live in the optimisers' registries (persistent buffers refreshed by ONE launch after each Adam
        # step, which IS captured below); the warm-up registered every operand the step uses.
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.out = gan_train_step(self.g, self.d, self.og, self.od, self.x, self.t, **self.kw)
        self.launches_per_step = _lib.launch_count - l0
        with torch.no_grad():
            for live, saved in snap:
                live.copy_(saved)
        optimizer_g._step, optimizer_d._step = steps0
        optimizer_g.flat_g.zero_()
        optimizer_d.flat_g.zero_()
        ops.bump_weight_epoch()
        optimizer_g.packs.refresh()        # the restored parameters' operands (the graph reads these buffers first thing)
        optimizer_d.packs.refresh()
        self._epoch_seen = ops._WEIGHT_EPOCH

    def load(self, input, target, non_blocking=True):
        self.x.copy_(input, non_blocking=non_blocking)
        self.t.copy_(target, non_blocking=non_blocking)

    def replay(self):
        if ops._WEIGHT_EPOCH != self._epoch_seen:
            # somebody changed the weights since the last replay (load_state_dict, a weight clamp, an eager step): the graph reads
            # the registries' buffers directly, so bring them up to date first
            self.og.packs.refresh()
            self.od.packs.refresh()
        self.og.sync_hyperparams()         # follow param_group changes (LR schedulers) made since the capture
        self.od.sync_hyperparams()
        self.graph.replay()
        self.og._step += 1
        self.od._step += 1
        ops.bump_weight_epoch()            # parameters changed behind autograd's back: drop eager packed-weight caches ...
        self.og.packs.restamp()            # ... except the registries' operands, which the replayed graph just refreshed
        self.od.packs.restamp()
        self._epoch_seen = ops._WEIGHT_EPOCH
        return self.out

    def metrics(self):
        """(iou, dice) of the last replayed step: device -> host read of the counts / leaf sums + the host halves."""
        finish_metrics(self.out)
        return self.out["iou"], self.out["dice"]

    def __call__(self, input, target):
        self.load(input, target)
        return self.replay()


# ----------------------------------------------------------------------------------------------
# the supervised trainer's loop body and the validation loop body
# ----------------------------------------------------------------------------------------------
def make_supervised_optimizer(model, config):
    """train.py:285-290 for `optimizer == 'Adam'`: Adam(lr, weight_decay) over every trainable parameter, as ONE
    flat-arena optimiser.  Parameters of modules that are constructed but never called (UNet_R_SS.sp_up1_3,
    archs.py:513) get no gradient in the reference, so torch's Adam skips them; they stay outside the arena."""
    from .archs import SubPixelConvolutionalBlock
    from .optim import FusedClampAdam
    if config.get("optimizer", "Adam") != "Adam":
        raise ops._lib.SsgError("make_supervised_optimizer: only config['optimizer'] == 'Adam' (config_v1.json) is built")
    idle = set()
    for m in model.modules():
        if isinstance(m, SubPixelConvolutionalBlock):
            idle.update(id(p) for p in m.parameters())
    params = [p for p in model.parameters() if p.requires_grad and id(p) not in idle]
    opt = FusedClampAdam(params, lr=float(config["lr"]), weight_decay=float(config.get("weight_decay", 0.0)))
    opt._idle_params = [p for p in model.parameters() if id(p) in idle]
    return opt


def _clamp_weights(model, optimizer, clip):
    """`for p in model.parameters(): p.data.clamp_(-clip, clip)` (train.py:111-112)."""
    if hasattr(optimizer, "clamp_weights"):
        optimizer.clamp_weights(clip)
        rest = getattr(optimizer, "_idle_params", [])
    else:
        rest = list(model.parameters())
    for p in rest:
        ops._lib.call("ssg_clamp_", p.data, p.numel(), float(clip))
    if rest:
        ops.bump_weight_epoch()


def _outputs_loss_metrics(config, model, criterion, input, target, with_metrics):
    """train.py:85-108 == train.py:155-175 == train_seg_gan.py:266-276: forward, loss, IoU / Dice."""
    num_class = int(config["num_classes"])
    iou = dice = None
    if config.get("deep_supervision"):
        outputs = model(input)
        loss = 0
        for output in outputs:
            loss = loss + criterion(output, target)
        loss = loss / len(outputs)
        if with_metrics:
            iou = iou_score(outputs[-1], target)
            dice = dice_coef(outputs[-1], target)
        return outputs[-1], loss, iou, dice
    output = ops.nan_to_zero(model(input))                                     # train.py:101
    if with_metrics:
        out_m = output[:, 1:num_class].detach().contiguous()                   # :102
        tar_m = target[:, 1:num_class].contiguous()                            # :103
        iou = iou_score(out_m, tar_m)                                          # :107
        dice = dice_coef(out_m, tar_m)                                         # :108
    loss = criterion(output, target)                                           # :105
    return output, loss, iou, dice


def supervised_train_step(config, model, criterion, optimizer, input, target, cnn_optimizer=None, epoch=0, with_metrics=True):
    """One iteration of the supervised trainer (train.py:81-120): forward, BCEDice (averaged over the outputs under deep
    supervision), metrics, the WEIGHT clamp to +-config['clip'] (placed, as in the reference, after the forward and before
    the backward), zero_grad / backward / step.  `model.train()` is the caller's job (train.py:73).
    Returns OrderedDict(loss, iou, dice, logits)."""
    output, loss, iou, dice = _outputs_loss_metrics(config, model, criterion, input, target, with_metrics)
    _clamp_weights(model, optimizer, float(config["clip"]))                    # :111-112
    optimizer.zero_grad()                                                      # :114
    loss.backward()                                                            # :115
    optimizer.step()                                                           # :116
    if cnn_optimizer is not None and epoch > 1:                                # :118-120
        cnn_optimizer.step()
    return OrderedDict([("loss", loss.detach()), ("iou", iou), ("dice", dice), ("logits", output.detach())])


def validate_step(config, model, criterion, input, target, with_metrics=True):
    """One iteration of `validate` (train.py:152-176, train_seg_gan.py:262-276) under no_grad; `model.eval()` is the
    caller's job (:146 / :259).  Returns OrderedDict(loss, iou, dice, logits)."""
    with torch.no_grad():
        output, loss, iou, dice = _outputs_loss_metrics(config, model, criterion, input, target, with_metrics)
    return OrderedDict([("loss", loss), ("iou", iou), ("dice", dice), ("logits", output)])


def validate(config, val_loader, model, criterion):
    """train.py:140-196 / train_seg_gan.py:253-294 without the progress bar: running averages weighted by batch size.
    `val_loader` yields the reference dataset's 5-tuples `(ori_img, input, target, targets, meta)` (dataset.py:144) or
    `(input, target)` pairs."""
    from .srgan_utils import AverageMeter
    avg = {"loss": AverageMeter(), "iou": AverageMeter(), "dice": AverageMeter()}
    model.eval()
    for batch in val_loader:
        input, target = (batch[1], batch[2]) if len(batch) >= 5 else (batch[0], batch[1])
        out = validate_step(config, model, criterion, input.cuda(), target.cuda())
        n = input.size(0)
        avg["loss"].update(out["loss"].item(), n)
        avg["iou"].update(out["iou"], n)
        avg["dice"].update(out["dice"], n)
    return OrderedDict([("loss", avg["loss"].avg), ("iou", avg["iou"].avg), ("dice", avg["dice"].avg)])
