"""Parity at the HEADLINE tile size (512 x 512) and on the 4-band SN7-shaped configuration, against fixtures written by the
unmodified reference (oracle/make_golden_headline.py).

How the logit tolerances are read.  The reference forward is a discontinuous function of its input: MaxPool argmax ->
MaxUnpool placement (archs.py:628-659) moves a value to another pixel whenever a near-tie flips.  The fixture records the
reference's OWN behaviour on this input: a relative input perturbation eps moves its fp32 logits by ~0.9 sqrt(eps) (2.9e-2 at
1e-3, 9.2e-3 at 1e-4, 2.8e-3 at 1e-5, 1.0e-3 at 1e-6), rounding only the input to bf16 moves them by 3.7e-2, and the
reference disagrees with itself by 2.2e-4 (one CPU thread), 6.6e-4 (ATen's native convolution) and 6.7e-4 (fp64 arithmetic)
on the SAME input.  So at this shape
  * the fp32 path is held to 3 x the reference's fp32-vs-fp64 self-deviation on the logits, 1e-4 on the losses;
  * the bf16 path is held to 1e-2 on the losses, to 1e-2 on every stage's output when the stage is fed the reference's
    own input (`test_stage_parity_teacher_forced`: no layer is off by more than bf16 rounding; measured 5.6e-3 .. 6.2e-3), and
    on the end-to-end logits to the reference's own response to a bf16-rounded input (x 2.5: a dozen storage sites instead of
    one; measured 7.5e-2 against 3.7e-2).  bf16 rounding alone flips 1.4e-3 .. 1.7e-3 of the MaxPool argmaxes of the
    reference's encoder outputs (same test), and each flip displaces one activation after MaxUnpool: sqrt(2 x 1.5e-3) = 5.5e-2.
"""
import functools
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SAMPLE = 2048


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(autouse=True)
def _reset():
    import ssunet_gan_b200 as ssg
    yield
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")


def _sample_idx(n):
    step = max(1, n // SAMPLE)
    return torch.arange(0, n, step)[:SAMPLE]


def _grad_errors(named_params, keys, norms, samples, clip=None, selferr=None):
    """Per-parameter rel-L2 of the strided gradient sample against the reference's (clamped like clip_gradient when the
    optimiser already ran).  Returns {key: (error, yardstick)}: the yardstick is the reference's OWN deviation for that parameter
    between two fp32 convolution implementations (oneDNN vs ATen native; `*_grad_selferr` in the fixture), floored at the
    median over all parameters.  Parameters whose reference gradient is noise even against itself (conv biases in front of a
    BatchNorm: mathematically zero) are skipped."""
    grads = {k: p.grad for k, p in named_params}
    med = float(np.median(selferr)) if selferr is not None else 0.0
    out = {}
    for i, (k, nrm, smp) in enumerate(zip(keys, norms, samples)):
        g = grads[str(k)].detach().reshape(-1).cpu()
        idx = _sample_idx(g.numel())
        want = torch.from_numpy(smp[:idx.numel()]).double()
        if clip is not None:
            want = want.clamp(-clip, clip)
        if nrm < 1e-6 or (selferr is not None and selferr[i] > 0.25):
            continue
        scale = max(float(want.norm()), 1e-30)
        out[str(k)] = (float((g[idx].double() - want).norm()) / scale, max(float(selferr[i]), med) if selferr is not None else None)
    return out


def _grad_summary(ge):
    errs = [e for e, _ in ge.values()]
    worst = max(ge.items(), key=lambda kv: kv[1][0] / (kv[1][1] or 1.0))
    return float(np.median(errs)), worst[0], worst[1][0], worst[1][1]


def _nets(O, dtype, impl, cin=3):
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import models_seg_gan
    ssg.set_compute_dtype(dtype)
    ssg.set_conv_impl(impl)
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": cin, "deep_supervision": False})
    g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, cin, prefix="net.")))
    d = models_seg_gan.Discriminator(3)
    d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3)))
    return g.cuda().train(), d.cuda().train()


@pytest.mark.parametrize("dtype,impl", [(torch.float32, "simt"), (torch.bfloat16, "auto")])
def test_headline_gan_step_vs_reference(golden_dir, dtype, impl):
    """One literal G+D iteration (train_seg_gan.py:188-233) on 2 x 3 x 512 x 512 against the unmodified reference."""
    import ssunet_oracle as O
    from ssunet_gan_b200 import optim, train_step
    z = np.load(os.path.join(golden_dir, "headline_gan_step_2x512.npz"))
    g, d = _nets(O, dtype, impl)
    og = optim.FusedClampAdam(g.parameters(), lr=2e-5)
    od = optim.FusedClampAdam(d.parameters(), lr=2e-5)
    x, t = O.synthetic_batch(2, 3, 512, 512, seed=1234, blobby=True)
    r = train_step.gan_train_step(g, d, og, od, x.cuda(), t.cuda())
    sc = z["scalars"]
    got = [float(r["loss"]), float(r["content"]), float(r["adv_g"]), float(r["adv_d"])]
    e_log = rel(r["logits"], z["logits"])
    ge = _grad_errors(g.named_parameters(), z["g_grad_keys"], z["g_grad_norm"], z["g_grad_sample"], 0.8, z["g_grad_selferr"])
    de = _grad_errors(d.named_parameters(), z["d_grad_keys"], z["d_grad_norm"], z["d_grad_sample"], 0.8, z["d_grad_selferr"])
    med_g, wk_g, we_g, wy_g = _grad_summary(ge)
    med_d, wk_d, we_d, wy_d = _grad_summary(de)
    ref_med_g, ref_med_d = float(np.median(z["g_grad_selferr"])), float(np.median(z["d_grad_selferr"]))
    print("\n[headline %s] scalars got %s want %s | logits rel-L2 %.3e (reference self: fp64 %.2e, bf16 input %.2e) | per-parameter "
          "gradient rel-L2: G median %.3e (reference vs itself %.3e), worst ratio %s %.3e vs %.3e | D median %.3e (%.3e), worst %s "
          "%.3e vs %.3e | iou %.6f/%.6f dice %.6f/%.6f"
          % (str(dtype).split(".")[-1], ["%.6f" % v for v in got], ["%.6f" % v for v in sc[:4]], e_log, float(z["self_fp64"]),
             float(z["sens_input_bf16"]), med_g, ref_med_g, wk_g, we_g, wy_g, med_d, ref_med_d, wk_d, we_d, wy_d,
             r["iou"], sc[4], float(r["dice"]), sc[5]))
    if dtype == torch.float32:
        assert e_log < 3 * float(z["self_fp64"]), e_log
        for a, b, tl in zip(got, sc[:4], (1e-4, 1e-4, 1e-3, 1e-3)):
            assert abs(a - b) < tl * abs(b), (got, sc)
        assert abs(r["iou"] - sc[4]) < 2e-4 and abs(float(r["dice"]) - sc[5]) < 2e-5
        # every parameter's gradient within 5 x what the reference's own two fp32 convolution back-ends differ by on it
        assert med_g < 2 * ref_med_g and med_d < 2 * ref_med_d, (med_g, ref_med_g, med_d, ref_med_d)
        bad = {k: v for k, v in list(ge.items()) + list(de.items()) if v[0] > 5 * v[1]}
        assert not bad, bad
    else:
        for a, b in zip(got, sc[:4]):
            assert abs(a - b) < 1e-2 * abs(b), (got, sc)            # north_star: bf16 within 1e-2 on the losses
        assert e_log < 2.5 * float(z["sens_input_bf16"]), e_log      # see the module docstring
        assert abs(r["iou"] - sc[4]) < 1e-2 and abs(float(r["dice"]) - sc[5]) < 1e-2
        # Gradients are more sensitive still than the logits (argmax / ReLU-mask flips change which paths carry gradient): the
        # reference's OWN per-parameter gradients move by a median of 0.24 (G) / 0.27 (D) rel-L2 when only its input is rounded to
        # bf16 (`*_grad_sens_bf16in` in the fixture).  The bf16 path (a dozen storage sites) is held to 2 x that on the median and
        # to 4 x the per-parameter figure (floored at the median) on every parameter.
        sg, sd_ = z["g_grad_sens_bf16in"], z["d_grad_sens_bf16in"]
        assert med_g < 2 * float(np.median(sg)) and med_d < 2 * float(np.median(sd_)), (med_g, med_d)
        for errs, keys, sens in ((ge, z["g_grad_keys"], sg), (de, z["d_grad_keys"], sd_)):
            yard = {str(k): max(float(v), float(np.median(sens))) for k, v in zip(keys, sens)}
            bad = {k: (v[0], yard[k]) for k, v in errs.items() if v[0] > 4 * yard[k]}
            assert not bad, bad


def test_headline_metrics_bit_exact_on_reference_logits(golden_dir):
    """IoU / Dice from this package's kernels on the REFERENCE's 2 x 3 x 512 x 512 logits (identical predicted masks):
    IoU bit-identical, Dice to the last float32 digit of numpy's pairwise sum (metrics.py:6-35)."""
    import ssunet_oracle as O
    from ssunet_gan_b200 import metrics
    z = np.load(os.path.join(golden_dir, "headline_gan_step_2x512.npz"))
    _, t = O.synthetic_batch(2, 3, 512, 512, seed=1234, blobby=True)
    lo = torch.from_numpy(z["logits"]).cuda()
    out_m, tar_m = lo[:, 1:3].contiguous(), t[:, 1:3].contiguous().cuda()
    assert metrics.iou_score(out_m, tar_m) == float(z["scalars"][4])
    assert abs(float(metrics.dice_coef(out_m, tar_m)) - float(z["scalars"][5])) < 2e-6


def _oracle_stages(O, sd, x):
    """The oracle forward (archs.py:623-671) stage by stage: [(block, spade, input, output)] plus the pooling indices."""
    import torch.nn.functional as F
    P = "net."
    rec = []

    def stage(c, s, t):
        y = O.spade(sd, P + s, O.basic_block(sd, P + c, t, True))
        rec.append((c, s, t, y))
        return y

    with torch.no_grad():
        e0 = stage("conv0_0", "SPADE0_0", x); p0, _ = F.max_pool2d(e0, 2, 2, return_indices=True)
        e1 = stage("conv1_0", "SPADE1_0", p0); p1, _ = F.max_pool2d(e1, 2, 2, return_indices=True)
        e2 = stage("conv2_0", "SPADE2_0", p1); p2, i2 = F.max_pool2d(e2, 2, 2, return_indices=True)
        e3 = stage("conv3_0", "SPADE3_0", p2); p3, i3 = F.max_pool2d(e3, 2, 2, return_indices=True)
        e4 = stage("conv4_0", "SPADE4_0", p3); p4, i4 = F.max_pool2d(e4, 2, 2, return_indices=True)
        e5 = F.conv2d(stage("conv5_0", "SPADE5_0", p4), sd[P + "conv_head5_0.weight"])
        d4 = F.conv2d(stage("conv4_1", "SPADE4_1", torch.cat([e4, F.max_unpool2d(e5, i4, 2, 2)], 1)), sd[P + "conv_head4_1.weight"])
        d3 = F.conv2d(stage("conv3_1", "SPADE3_1", torch.cat([e3, F.max_unpool2d(d4, i3, 2, 2)], 1)), sd[P + "conv_head3_1.weight"])
        d2 = stage("conv2_1", "SPADE2_1", torch.cat([e2, F.max_unpool2d(d3, i2, 2, 2)], 1))
        d1 = stage("conv1_1", "SPADE1_1", torch.cat([e1, O._up(d2)], 1))
        d0 = stage("conv0_1", "SPADE0_1", torch.cat([e0, O._up(d1)], 1))
        logits = F.conv2d(d0, sd[P + "final.weight"], sd[P + "final.bias"])
    return rec, logits, {"e2": (e2, i2), "e3": (e3, i3), "e4": (e4, i4)}


def test_stage_parity_teacher_forced(golden_dir):
    """"Which layer breaks it?"  None: every BasicBlock + SPADE stage of the bf16 tensor-core path, fed the fp32 reference's
    own stage input at 2 x 3 x 512 x 512, reproduces the reference's stage output within 1e-2 rel-L2 (north_star's bf16
    bound).  The end-to-end logit gap is the composition: bf16 rounding flips a small fraction of MaxPool argmaxes (measured
    below on the reference's own tensors) and MaxUnpool turns each flip into a displaced activation."""
    import ssunet_oracle as O
    from ssunet_gan_b200 import ops
    z = np.load(os.path.join(golden_dir, "headline_gan_step_2x512.npz"))
    sd = O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net."))
    x, _ = O.synthetic_batch(2, 3, 512, 512, seed=1234, blobby=True)
    rec, logits, pooled = _oracle_stages(O, dict(sd), x)
    assert rel(logits, z["logits"]) < 3 * float(z["self_1thread"]) + 1e-6      # the oracle wiring above IS the reference forward
    g, _ = _nets(O, torch.bfloat16, "auto")
    net = g.net
    worst = 0.0
    with torch.no_grad():
        for c, s, xin, want in rec:
            # thin inputs (the 3-band image) enter channel-padded, exactly as UNet_R_SS_v2.forward stores them
            blk_out = getattr(net, c)(ops.to_nhwc(xin.cuda(), pad_channels=xin.shape[1] < 8))
            y = getattr(net, s)(blk_out, blk_out)
            e = rel(ops.to_nchw_f32(y), want)
            print("stage %-8s + %-9s in %-22s rel-L2 %.3e" % (c, s, tuple(xin.shape), e))
            worst = max(worst, e)
        # argmax flips caused by bf16 rounding alone, on the reference's encoder outputs whose indices feed MaxUnpool
        for name, (e_ref, idx) in pooled.items():
            _, code = ops.max_pool2x2(ops.to_nhwc(e_ref.cuda()))
            n, c_, h, w = e_ref.shape
            oy = torch.arange(h // 2).view(1, 1, -1, 1)
            ox = torch.arange(w // 2).view(1, 1, 1, -1)
            want_code = ((idx // w) - 2 * oy) * 2 + ((idx % w) - 2 * ox)          # 2-bit position inside the 2x2 window
            got_code = code.permute(0, 3, 1, 2).cpu().long()
            flips = float((got_code != want_code).float().mean())
            print("MaxPool argmax flips from bf16 rounding of %s: %.3e of the windows" % (name, flips))
            assert flips < 2e-2
    assert worst < 1e-2, worst


@pytest.mark.parametrize("dtype,impl", [(torch.float32, "simt"), (torch.bfloat16, "auto")])
def test_sn7_four_band_training_step(golden_dir, dtype, impl):
    """BASELINE configs[3] layout: Generator(input_channels=4) forward + BCEDice + backward + clip + Adam on 2 x 4 x 64 x 64
    against the unmodified reference (archs.py:576: Cin is a constructor argument)."""
    import ssunet_oracle as O
    from ssunet_gan_b200 import ops, optim
    from ssunet_gan_b200.srgan_utils import clip_gradient
    z = np.load(os.path.join(golden_dir, "sn7_train_step_2x4x64.npz"))
    g, _ = _nets(O, dtype, impl, cin=4)
    opt = optim.FusedClampAdam(g.parameters(), lr=2e-5)
    x, t = O.synthetic_batch(2, 4, 64, 64, seed=4321, blobby=True)
    out = g(x.cuda())
    loss = ops.seg_losses(out, t.cuda())[0]
    opt.zero_grad()
    loss.backward()
    ge = {k: v[0] for k, v in _grad_errors(g.named_parameters(), z["grad_keys"], z["grad_norm"], z["grad_sample"]).items()}
    clip_gradient(opt, 0.8)
    opt.step()
    torch.cuda.synchronize()
    e_log = rel(out, z["logits"])
    worst = max(ge.items(), key=lambda kv: kv[1])
    med = float(np.median(list(ge.values())))
    print("\n[sn7 %s] logits rel-L2 %.3e loss %.6f/%.6f grads median %.3e worst %.3e (%s)"
          % (str(dtype).split(".")[-1], e_log, float(loss), float(z["loss"]), med, worst[1], worst[0]))
    sdg = g.state_dict()
    if dtype == torch.float32:
        assert e_log < 1e-4 and abs(float(loss) - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
        assert med < 2e-3 and worst[1] < 5e-2, (med, worst)
        for i, k in enumerate(z["upd_keys"]):
            # one clamped Adam step moves every element by <= lr; elements whose gradient is rounding noise may move either way
            np.testing.assert_allclose(sdg[str(k)].cpu().numpy(), z["upd_%d" % i], rtol=0, atol=4.1e-5)
        w = sdg["net.conv0_0.conv1.weight"].cpu().numpy()
        assert w.shape[1] == 4
        assert np.mean(np.abs(w - z["upd_0"]) < 1e-7) > 0.98          # and almost all of them land on the reference's value
    else:
        assert abs(float(loss) - float(z["loss"])) < 1e-2 * abs(float(z["loss"]))
        assert e_log < 0.12          # 2x2 bottleneck, BatchNorm over 8 samples (tests/test_gpu_modules.py explains)
        assert med < 0.6, med        # per-parameter gradient rel-L2; the headline-shape test above carries the yardstick for this
