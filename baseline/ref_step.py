"""The reference's own CPU implementation of the G+D training step, driven through its own modules (bench.py's
`--impl reference` arm and `cpu_baseline` leg; kind = "reference").

`baseline/_ref/scripts` is a verbatim copy of the reference's scripts (baseline/make_ref.py).  `train_seg_gan.train()` itself
moves every batch to the GPU (`.cuda()`, train_seg_gan.py:183-184), so the loop body (train_seg_gan.py:188-233) is restated
here line by line WITHOUT those two calls, driving the reference's unmodified `Generator`, `Discriminator`, `BCEDiceLoss`,
`iou_score`, `dice_coef`, `clip_gradient` and torch.optim.Adam exactly as `main()` builds them (train_seg_gan.py:448-472).
Nothing from this repository's package is on that path.
"""
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "scripts")


def available():
    return os.path.isfile(os.path.join(REF, "train_seg_gan.py"))


def _stub_missing_imports():
    """albumentations / tensorboardX / torchsummary are imported at module top by the reference's trainers
    (train_seg_gan.py:15-20) and are not installed in this image; the loop body never calls them."""
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            raise RuntimeError("stubbed dependency called")

    try:
        import albumentations  # noqa: F401
    except Exception:
        mod("albumentations", Compose=_Any, OneOf=_Any, Resize=_Any, Normalize=_Any, Flip=_Any, RandomRotate90=_Any)
        mod("albumentations.augmentations", transforms=mod("albumentations.augmentations.transforms", **{
            k: _Any for k in ("Flip", "Normalize", "Resize", "RandomRotate90", "HueSaturationValue", "RandomBrightness",
                              "RandomContrast", "RandomBrightnessContrast", "Rotate")}))
        mod("albumentations.core", composition=mod("albumentations.core.composition", Compose=_Any, OneOf=_Any))
    try:
        import tensorboardX  # noqa: F401
    except Exception:
        mod("tensorboardX", SummaryWriter=_Any)
    try:
        import torchsummary  # noqa: F401
    except Exception:
        mod("torchsummary", summary=lambda *a, **k: None)


def load():
    """Import the reference's modules from baseline/_ref/scripts; returns a namespace of the ones the step uses."""
    if not available():
        raise RuntimeError("baseline/_ref/scripts is missing: run `python baseline/make_ref.py` in the build container")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _stub_missing_imports()
    import warnings
    warnings.filterwarnings("ignore")
    import models_seg_gan, losses, metrics, srgan_utils, batchnorm  # noqa: E401  (the reference's own files)
    ns = types.SimpleNamespace(models_seg_gan=models_seg_gan, losses=losses, metrics=metrics, srgan_utils=srgan_utils,
                               batchnorm=batchnorm)
    assert os.path.realpath(models_seg_gan.__file__).startswith(os.path.realpath(REF)), models_seg_gan.__file__
    return ns


class ReferenceStep:
    """Generator / Discriminator / optimisers / criteria as train_seg_gan.main() builds them (seed 41, :35-36), and the loop body."""

    def __init__(self, input_channels=3, num_classes=3, lr=2e-5):
        import torch
        self.torch = torch
        self.R = load()
        torch.manual_seed(41)                                                           # train_seg_gan.py:35-36
        cfg = {"arch": "UNet_R_SS_v2", "num_classes": num_classes, "input_channels": input_channels, "deep_supervision": False}
        self.generator = self.R.models_seg_gan.Generator(cfg)                            # :448
        self.optimizer_g = torch.optim.Adam(params=filter(lambda p: p.requires_grad, self.generator.parameters()), lr=lr)   # :452
        self.discriminator = self.R.models_seg_gan.Discriminator(num_classes=num_classes)               # :463-466
        self.optimizer_d = torch.optim.Adam(params=filter(lambda p: p.requires_grad, self.discriminator.parameters()), lr=lr)  # :468
        self.criterion = self.R.losses.BCEDiceLoss()                                      # :342
        self.content_loss_criterion = torch.nn.MSELoss()                                  # :471
        self.adversarial_loss_criterion = torch.nn.BCEWithLogitsLoss()                    # :472
        self.num_classes = num_classes
        self.generator.train()
        self.discriminator.train()                                                        # :176-177

    def state_dicts(self):
        return ({k: v.detach().clone() for k, v in self.generator.state_dict().items()},
                {k: v.detach().clone() for k, v in self.discriminator.state_dict().items()})

    def step(self, input, target, alpa=1e-4, beta=1e-3, grad_clip=0.8):
        """train_seg_gan.py:188-233 verbatim, minus the two `.cuda()` calls of :183-184."""
        torch = self.torch
        R = self.R
        num_class = self.num_classes
        generator, discriminator = self.generator, self.discriminator
        optimizer_g, optimizer_d = self.optimizer_g, self.optimizer_d
        generator_output = generator(input)                                                  # :188
        generator_output[torch.isnan(generator_output)] = 0                                 # :190
        out_m = generator_output[:, 1:num_class, :, :].clone()                              # :191
        tar_m = target[:, 1:num_class, :, :].clone()                                        # :192
        loss = self.criterion(generator_output, target)                                     # :194
        content_loss = self.content_loss_criterion(generator_output, target)                # :195
        iou = R.metrics.iou_score(out_m, tar_m)                                             # :197
        dice = R.metrics.dice_coef(out_m, tar_m)                                            # :198
        seg_discriminated = discriminator(generator_output)                                 # :202
        adversarial_loss = self.adversarial_loss_criterion(seg_discriminated, torch.ones_like(seg_discriminated))   # :204
        perceptual_loss = loss + alpa * content_loss + beta * adversarial_loss              # :205
        optimizer_g.zero_grad()                                                             # :207
        perceptual_loss.backward()                                                          # :208
        if grad_clip is not None:
            R.srgan_utils.clip_gradient(optimizer_g, grad_clip)                             # :211-212
        optimizer_g.step()                                                                  # :215
        adv_g = float(adversarial_loss.item())
        hr_discriminated = discriminator(target)                                            # :217
        sr_discriminated = discriminator(generator_output.detach())                         # :218
        adversarial_loss = self.adversarial_loss_criterion(sr_discriminated, torch.zeros_like(sr_discriminated)) + \
            self.adversarial_loss_criterion(hr_discriminated, torch.ones_like(hr_discriminated))    # :221-222
        optimizer_d.zero_grad()                                                             # :225
        adversarial_loss.backward()                                                         # :226
        if grad_clip is not None:
            R.srgan_utils.clip_gradient(optimizer_d, grad_clip)                             # :229-230
        optimizer_d.step()                                                                  # :233
        return {"loss": float(loss.item()), "content": float(content_loss.item()), "adv_g": adv_g,
                "adv_d": float(adversarial_loss.item()), "iou": float(iou), "dice": float(dice),
                "logits": generator_output.detach()}


def timed_steps(batch, size, input_channels=3, steps=1, warmup=0, seed=1234, make_batch=None):
    """Run `warmup + steps` reference iterations on synthetic batches; returns (seconds per timed step, first step's result,
    the initial state_dicts).  All host threads."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceStep(input_channels=input_channels)
    init = ref.state_dicts()
    times, first = [], None
    for it in range(warmup + steps):
        x, t = make_batch(batch, input_channels, size, size, seed=seed + it)
        t0 = time.perf_counter()
        r = ref.step(x, t)
        dt = time.perf_counter() - t0
        if it == 0:
            first = r
        if it >= warmup:
            times.append(dt)
    return times, first, init
