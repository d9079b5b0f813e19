"""Uninitialised-read hunt: every torch.empty / empty_like made while the package runs is filled with NaN (floats) or
0x7f bytes first; a kernel that reads an element it (or an earlier kernel) never wrote turns the outputs into NaN."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
_e, _el = torch.empty, torch.empty_like


def _poison(t):
    if t.is_cuda and t.numel():
        if t.is_floating_point():
            t.fill_(float("nan"))
        else:
            t.fill_(127)
    return t


torch.empty = lambda *a, **k: _poison(_e(*a, **k))
torch.empty_like = lambda *a, **k: _poison(_el(*a, **k))
import ssunet_oracle as O
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import models_seg_gan, optim, train_step, ops


def check(tag, **ts):
    for k, v in ts.items():
        if v is None:
            continue
        bad = int(torch.isnan(v.float()).sum())
        print("%-28s %-10s nan=%d / %d" % (tag, k, bad, v.numel()), flush=True)


for dt, impl in ((torch.float32, "simt"), (torch.bfloat16, "auto")):
    ssg.set_compute_dtype(dt); ssg.set_conv_impl(impl)
    for (b, s) in ((2, 64), (1, 96), (2, 80)):
        g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
        g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
        d = models_seg_gan.Discriminator(3)
        d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3)))
        g.cuda().train(); d.cuda().train()
        x, t = O.synthetic_batch(b, 3, s, s, seed=1234)
        x, t = x.cuda(), t.cuda()
        out = g(x)
        loss = ops.seg_losses(out, t)[0]
        loss.backward()
        gr = torch.cat([p.grad.reshape(-1) for p in g.parameters() if p.grad is not None])
        check("G %s %dx%d" % (str(dt)[6:], b, s), logits=out, loss=loss, grads=gr)
        og = optim.FusedClampAdam(g.parameters(), lr=2e-5); od = optim.FusedClampAdam(d.parameters(), lr=2e-5)
        r = train_step.gan_train_step(g, d, og, od, x, t, with_metrics=False)
        check("step %s %dx%d" % (str(dt)[6:], b, s), logits=r["logits"], adv_d=r["adv_d"], pg=og.flat_p, pd=od.flat_p)
