"""Fixtures at the HEADLINE tile size (512 x 512) and for the 4-band SN7-shaped configuration, written by the UNMODIFIED
reference (build container only; the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_headline.py

  tests/golden/headline_gan_step_2x512.npz   one literal G+D iteration (train_seg_gan.py:188-233) on 2 x 3 x 512 x 512:
        logits (full, fp32), the six scalars, a strided sample (<= 2048 elements) + the L2 norm of EVERY parameter gradient
        of both networks (before clip_gradient), and the reference's own sensitivity figures (see below).
  tests/golden/sn7_train_step_2x4x64.npz     Generator(input_channels=4) forward + BCEDice + backward + clip + Adam step
        on 2 x 4 x 64 x 64 (preprocess_SN7 layout: 4-band tiles), logits / loss / gradient samples / updated parameters.

Sensitivity figures: the reference forward is a DIScontinuous function of its input (MaxPool argmax -> MaxUnpool placement,
archs.py:628-659: a near-tie that flips moves a value to another pixel), so a relative perturbation eps of the input moves the
fp32 reference's own logits by ~ sqrt(eps), not eps.  `sens_*` record rel-L2(logits(x'), logits(x)) of the reference for
x' = bf16(x) and x' = x (1 + eps n), eps = 1e-3 .. 1e-6 -- the yardstick the bf16 tests are read against.  `self_*` record
the reference against itself on the SAME input: one CPU thread instead of all (another summation order), ATen's native
convolution instead of oneDNN, and fp64 arithmetic -- the floor any fp32 implementation can be compared at.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/scripts")
warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True

import ssunet_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(os.cpu_count() or 8)
SAMPLE = 2048


def sample_idx(n):
    """The fixed strided sample of a flattened parameter gradient (the tests index with the same rule)."""
    step = max(1, n // SAMPLE)
    return torch.arange(0, n, step)[:SAMPLE]


def grad_record(named_params):
    keys, norms, samples = [], [], []
    for k, p in named_params:
        if p.grad is None:
            continue
        g = p.grad.detach().reshape(-1)
        keys.append(k)
        norms.append(float(g.double().norm()))
        s = torch.zeros(SAMPLE)
        idx = sample_idx(g.numel())
        s[:idx.numel()] = g[idx]
        samples.append(s.numpy())
    return np.array(keys), np.array(norms, dtype=np.float64), np.stack(samples).astype(np.float32)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def headline():
    import models_seg_gan, losses, metrics, srgan_utils  # noqa
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
    d = models_seg_gan.Discriminator(3)
    g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
    d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3)))
    g.train(); d.train()
    opt_g = torch.optim.Adam(filter(lambda p: p.requires_grad, g.parameters()), lr=2e-5)
    opt_d = torch.optim.Adam(filter(lambda p: p.requires_grad, d.parameters()), lr=2e-5)
    crit = losses.BCEDiceLoss()
    adv_c = torch.nn.BCEWithLogitsLoss()
    con_c = torch.nn.MSELoss()
    inp, tar = O.synthetic_batch(2, 3, 512, 512, seed=1234, blobby=True)
    # ---- the reference's own sensitivity to input perturbations (train-mode forward; BN buffers restored afterwards) ----
    bufs = {k: v.clone() for k, v in g.state_dict().items()}
    sens = {}
    with torch.no_grad():
        base = g(inp)
        sens["sens_input_bf16"] = rel(g(inp.bfloat16().float()), base)
        gen = torch.Generator().manual_seed(5)
        noise = torch.randn(inp.shape, generator=gen)
        for eps in (1e-3, 1e-4, 1e-5, 1e-6):
            sens["sens_eps_%g" % eps] = rel(g(inp * (1 + eps * noise)), base)
        # the reference against ITSELF on the same input: other summation orders / exact arithmetic
        nt = torch.get_num_threads()
        torch.set_num_threads(1)
        sens["self_1thread"] = rel(g(inp), base)
        torch.set_num_threads(nt)
        with torch.backends.mkldnn.flags(enabled=False):
            sens["self_native_conv"] = rel(g(inp), base)
        import copy
        sens["self_fp64"] = rel(copy.deepcopy(g).double()(inp.double()).float(), base)
    g.load_state_dict(bufs)
    print({k: "%.3e" % v for k, v in sens.items()}, flush=True)
    # ---- literal restatement of train_seg_gan.py:188-233 driving the reference modules ----
    go = g(inp)
    go[torch.isnan(go)] = 0
    out_m = go[:, 1:3, :, :].clone()
    tar_m = tar[:, 1:3, :, :].clone()
    loss = crit(go, tar)
    content = con_c(go, tar)
    iou = metrics.iou_score(out_m, tar_m)
    dice = metrics.dice_coef(out_m, tar_m)
    sdisc = d(go)
    adv = adv_c(sdisc, torch.ones_like(sdisc))
    perceptual = loss + 1e-4 * content + 1e-3 * adv
    opt_g.zero_grad()
    perceptual.backward()
    gk, gn, gs = grad_record(g.named_parameters())
    srgan_utils.clip_gradient(opt_g, 0.8)
    opt_g.step()
    hr = d(tar)
    sr = d(go.detach())
    advd = adv_c(sr, torch.zeros_like(sr)) + adv_c(hr, torch.ones_like(hr))
    opt_d.zero_grad()
    advd.backward()
    dk, dn, ds = grad_record(d.named_parameters())
    srgan_utils.clip_gradient(opt_d, 0.8)
    opt_d.step()
    scalars = np.array([loss.item(), content.item(), adv.item(), advd.item(), float(iou), float(dice)], dtype=np.float64)
    print("scalars", scalars, flush=True)
    # ---- the reference's gradients against ITSELF: the same iteration with ATen's native convolution instead of oneDNN
    # (another fp32 summation order).  Per parameter: rel-L2 of the strided sample -- what "equal up to fp32 rounding"
    # means for each gradient of this discontinuous network; the GPU tests hold each parameter to a multiple of it.
    g2 = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
    d2 = models_seg_gan.Discriminator(3)
    g2.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
    d2.load_state_dict(O.portable_state_dict(O.discriminator_spec(3)))
    g2.train(); d2.train()
    with torch.backends.mkldnn.flags(enabled=False):
        go2 = g2(inp)
        l2 = crit(go2, tar) + 1e-4 * con_c(go2, tar) + 1e-3 * adv_c(d2(go2), torch.ones(2, 1))
        l2.backward()
        gk2, gn2, gs2 = grad_record(g2.named_parameters())
        for p_ in d2.parameters():
            p_.grad = None
        a2 = adv_c(d2(go2.detach()), torch.zeros(2, 1)) + adv_c(d2(tar), torch.ones(2, 1))
        a2.backward()
        dk2, dn2, ds2 = grad_record(d2.named_parameters())
    assert list(gk2) == list(gk) and list(dk2) == list(dk)
    # ---- and against itself when only its INPUT (and the masks' consumer: nothing else) is rounded to bf16: the gradient
    # counterpart of `sens_input_bf16`, the yardstick for the bf16 tensor-core path's per-parameter gradients
    g3 = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
    d3 = models_seg_gan.Discriminator(3)
    g3.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
    d3.load_state_dict(O.portable_state_dict(O.discriminator_spec(3)))
    g3.train(); d3.train()
    go3 = g3(inp.bfloat16().float())
    l3 = crit(go3, tar) + 1e-4 * con_c(go3, tar) + 1e-3 * adv_c(d3(go3), torch.ones(2, 1))
    l3.backward()
    gk3, gn3, gs3 = grad_record(g3.named_parameters())
    for p_ in d3.parameters():
        p_.grad = None
    a3 = adv_c(d3(go3.detach()), torch.zeros(2, 1)) + adv_c(d3(tar), torch.ones(2, 1))
    a3.backward()
    dk3, dn3, ds3 = grad_record(d3.named_parameters())

    def selferr(a, b, norms):
        out = []
        for sa, sb, nrm in zip(a, b, norms):
            den = max(float(np.linalg.norm(sa.astype(np.float64))), 1e-30)
            out.append(float(np.linalg.norm(sa.astype(np.float64) - sb.astype(np.float64))) / den)
        return np.array(out)
    g_self, d_self = selferr(gs, gs2, gn), selferr(ds, ds2, dn)
    g_b16, d_b16 = selferr(gs, gs3, gn), selferr(ds, ds3, dn)
    print("gradient self-deviation (oneDNN vs native conv): G median %.3e max %.3e | D median %.3e max %.3e"
          % (np.median(g_self), g_self.max(), np.median(d_self), d_self.max()), flush=True)
    print("gradient deviation under a bf16-rounded input: G median %.3e max %.3e | D median %.3e max %.3e"
          % (np.median(g_b16), g_b16.max(), np.median(d_b16), d_b16.max()), flush=True)
    np.savez_compressed(os.path.join(OUT, "headline_gan_step_2x512.npz"), logits=go.detach().numpy(), scalars=scalars,
                        sr_logit=sr.detach().numpy(), hr_logit=hr.detach().numpy(),
                        g_grad_keys=gk, g_grad_norm=gn, g_grad_sample=gs, d_grad_keys=dk, d_grad_norm=dn, d_grad_sample=ds,
                        g_grad_selferr=g_self, d_grad_selferr=d_self, g_grad_sens_bf16in=g_b16, d_grad_sens_bf16in=d_b16,
                        **{k: np.float64(v) for k, v in sens.items()})


def sn7():
    import models_seg_gan, losses, srgan_utils  # noqa
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 4, "deep_supervision": False})
    spec = O.unet_r_ss_v2_spec(3, 4, prefix="net.")
    g.load_state_dict(O.portable_state_dict(spec))
    g.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, g.parameters()), lr=2e-5)
    x, t = O.synthetic_batch(2, 4, 64, 64, seed=4321, blobby=True)
    out = g(x)
    loss = losses.BCEDiceLoss()(out, t)
    opt.zero_grad()
    loss.backward()
    gk, gn, gs = grad_record(g.named_parameters())
    srgan_utils.clip_gradient(opt, 0.8)
    opt.step()
    sd = g.state_dict()
    first = {k: sd[k].numpy() for k in ("net.conv0_0.conv1.weight", "net.conv0_0.shortcut.0.weight", "net.final.weight", "net.final.bias")}
    np.savez_compressed(os.path.join(OUT, "sn7_train_step_2x4x64.npz"), logits=out.detach().numpy(), loss=np.float64(loss.item()),
                        grad_keys=gk, grad_norm=gn, grad_sample=gs,
                        upd_keys=np.array(list(first)), **{"upd_%d" % i: v for i, v in enumerate(first.values())})
    print("sn7 loss", loss.item(), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["headline", "sn7"]
    if "sn7" in which:
        sn7()
    if "headline" in which:
        headline()
