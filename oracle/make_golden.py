"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Imports /root/reference/scripts (read-only), loads the portable weights defined in
oracle/ssunet_oracle.py into the reference nn.Modules, drives them on the portable synthetic
batches and stores inputs' seeds + outputs as small fixtures.  The GPU box has no
/root/reference; tests only read the committed fixtures.
"""
import os
import sys
import json
import zlib
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
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


def csum(t):
    t = t.detach().double()
    return np.array([float(t.sum()), float(t.abs().sum()), float((t * t).sum())])


def main():
    import archs, models_seg_gan, losses, metrics, batchnorm, spectral_norm, srgan_utils, xresidualblock  # noqa

    # ---------------- state_dict layout + default-init checksums (seed 41, G then D) -----------
    torch.manual_seed(41)
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3,
                                  "deep_supervision": False})
    d = models_seg_gan.Discriminator(3)
    layout = {
        "generator": [[k, list(v.shape)] for k, v in g.state_dict().items()],
        "discriminator": [[k, list(v.shape)] for k, v in d.state_dict().items()],
        "init_seed41_generator": {k: csum(v).tolist() for k, v in g.state_dict().items() if v.is_floating_point()},
        "init_seed41_discriminator": {k: csum(v).tolist() for k, v in d.state_dict().items() if v.is_floating_point()},
    }
    with open(os.path.join(OUT, "state_layout.json"), "w") as f:
        json.dump(layout, f)

    # ---------------- generator forward/backward, portable weights, 2x3x64x64 -------------------
    spec_g = O.unet_r_ss_v2_spec(3, 3, prefix="net.")
    sd = O.portable_state_dict(spec_g)
    g.load_state_dict(sd)
    g.train()
    x, t = O.synthetic_batch(2, 3, 64, 64, seed=1234)
    out = g(x)
    crit = losses.BCEDiceLoss()
    loss = crit(out, t)
    loss.backward()
    gr = {k: csum(p.grad) for k, p in g.named_parameters() if p.grad is not None}
    np.savez_compressed(
        os.path.join(OUT, "generator_fwd_bwd_2x64.npz"),
        logits=out.detach().numpy(), loss=np.float64(loss.item()),
        iou=np.float64(metrics.iou_score(out[:, 1:], t[:, 1:])),
        dice=np.float32(metrics.dice_coef(out[:, 1:].clone(), t[:, 1:].clone())),
        grad_keys=np.array(list(gr.keys())), grad_csum=np.stack(list(gr.values())),
        bn_running_mean=g.state_dict()["net.conv0_0.bn1.running_mean"].numpy(),
        bn_running_var=g.state_dict()["net.conv2_1.bn2.running_var"].numpy(),
    )
    # eval-mode forward (config 5 semantics), different input seed
    g.load_state_dict(O.portable_state_dict(spec_g))   # fresh running stats
    g.eval()
    xe, _ = O.synthetic_batch(1, 3, 96, 96, seed=77)
    with torch.no_grad():
        oe = g(xe)
    np.savez_compressed(os.path.join(OUT, "generator_eval_1x96.npz"), logits=oe.numpy())

    # ---------------- discriminator forward/backward -------------------------------------------
    sd_d = O.portable_state_dict(O.discriminator_spec(3))
    d.load_state_dict(sd_d)
    d.train()
    xd, _ = O.synthetic_batch(3, 3, 96, 96, seed=5)
    xd.requires_grad_(True)
    lo = d(xd)
    l = torch.nn.functional.binary_cross_entropy_with_logits(lo, torch.ones_like(lo))
    l.backward()
    gd = {k: csum(p.grad) for k, p in d.named_parameters()}
    np.savez_compressed(os.path.join(OUT, "discriminator_fwd_bwd_3x96.npz"), logit=lo.detach().numpy(),
                        loss=np.float64(l.item()), dx=xd.grad.numpy(),
                        grad_keys=np.array(list(gd.keys())), grad_csum=np.stack(list(gd.values())))

    # ---------------- full G+D loop body, 2 iterations, 2x3x64x64 ------------------------------
    g.load_state_dict(O.portable_state_dict(spec_g)); g.train()
    d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.train()
    opt_g = torch.optim.Adam(filter(lambda p: p.requires_grad, g.parameters()), lr=2e-5)
    opt_d = torch.optim.Adam(filter(lambda p: p.requires_grad, d.parameters()), lr=2e-5)
    adv_c = torch.nn.BCEWithLogitsLoss()
    con_c = torch.nn.MSELoss()
    rec = {}
    for it in range(2):
        inp, tar = O.synthetic_batch(2, 3, 64, 64, seed=1234 + it, blobby=(it == 1))
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
        srgan_utils.clip_gradient(opt_g, 0.8)
        opt_g.step()
        hr = d(tar)
        sr = d(go.detach())
        advd = adv_c(sr, torch.zeros_like(sr)) + adv_c(hr, torch.ones_like(hr))
        opt_d.zero_grad()
        advd.backward()
        srgan_utils.clip_gradient(opt_d, 0.8)
        opt_d.step()
        rec["it%d_scalars" % it] = np.array([loss.item(), content.item(), adv.item(), advd.item(),
                                             float(iou), float(dice)], dtype=np.float64)
        rec["it%d_logits" % it] = go.detach().numpy()
    sdg = g.state_dict(); sdd = d.state_dict()
    rec["g_keys"] = np.array([k for k in sdg]); rec["g_csum"] = np.stack([csum(v) for v in sdg.values()])
    rec["d_keys"] = np.array([k for k in sdd]); rec["d_csum"] = np.stack([csum(v) for v in sdd.values()])
    rec["final_weight"] = sdg["net.final.weight"].numpy()
    rec["d_fc2_weight"] = sdd["fc2.weight"].numpy()
    np.savez_compressed(os.path.join(OUT, "gan_step_2it_2x64.npz"), **rec)

    # ---------------- SyncBN arithmetic: parallel-mode formulas on a sharded batch -------------
    # Parallel mode needs >= 2 CUDA devices (ReduceAddCoalesced/Broadcast); here we call the
    # reference's own _compute_mean_std on summed shard statistics (batchnorm.py:115-127).
    bn = batchnorm.SynchronizedBatchNorm2d(8)
    with torch.no_grad():
        bn.weight.copy_(O.portable_tensor("sbn.weight", (8,))); bn.bias.copy_(O.portable_tensor("sbn.bias", (8,)))
    gg = torch.Generator().manual_seed(9)
    xs = torch.randn(6, 8, 5, 7, generator=gg) * 2 + 0.5
    shards = xs.chunk(3, 0)
    s = sum(batchnorm._sum_ft(sh.reshape(sh.size(0), 8, -1)) for sh in shards)
    ss = sum(batchnorm._sum_ft(sh.reshape(sh.size(0), 8, -1) ** 2) for sh in shards)
    mean, inv_std = bn._compute_mean_std(s, ss, 6 * 35)
    y = (xs.reshape(6, 8, -1) - mean.view(1, -1, 1)) * (inv_std * bn.weight).view(1, -1, 1) + bn.bias.view(1, -1, 1)
    np.savez_compressed(os.path.join(OUT, "syncbn_3shards.npz"), x=xs.numpy(), y=y.detach().reshape(xs.shape).numpy(),
                        mean=mean.numpy(), inv_std=inv_std.numpy(),
                        running_mean=bn.running_mean.numpy(), running_var=bn.running_var.numpy())

    # ---------------- spectral norm (vendored spectral_norm.py) --------------------------------
    conv = torch.nn.Conv2d(6, 10, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(O.portable_tensor("sn.weight", (10, 6, 3, 3)))
    torch.manual_seed(3)
    spectral_norm.spectral_norm(conv)
    u0 = conv.weight_u.clone(); v0 = conv.weight_v.clone()
    conv.train()
    xi = torch.randn(2, 6, 8, 8, generator=torch.Generator().manual_seed(4))
    y1 = conv(xi)
    w1 = conv.weight.detach().clone(); u1 = conv.weight_u.clone(); v1 = conv.weight_v.clone()
    y1.sum().backward()
    np.savez_compressed(os.path.join(OUT, "spectral_norm_conv.npz"), u0=u0.numpy(), v0=v0.numpy(),
                        w_orig=conv.weight_orig.detach().numpy(), w1=w1.numpy(), u1=u1.numpy(), v1=v1.numpy(),
                        x=xi.numpy(), y1=y1.detach().numpy(), gw_orig=conv.weight_orig.grad.numpy(),
                        sd_keys=np.array(list(conv.state_dict().keys())))

    # ---------------- metrics & loss on hand-made cases ----------------------------------------
    gg = torch.Generator().manual_seed(11)
    lg = torch.randn(3, 2, 33, 47, generator=gg) * 3
    tg = (torch.rand(3, 2, 33, 47, generator=gg) > 0.6).float()
    hard = torch.where(lg > 0, torch.full_like(lg, 200.0), torch.full_like(lg, -200.0))
    l3 = torch.randn(3, 3, 33, 47, generator=gg) * 3
    t3 = (torch.rand(3, 3, 33, 47, generator=gg) > 0.5).float()
    np.savez_compressed(
        os.path.join(OUT, "metrics_loss.npz"), logits=lg.numpy(), target=tg.numpy(),
        iou=np.float64(metrics.iou_score(lg, tg)), dice=np.float32(metrics.dice_coef(lg, tg)),
        iou_hard=np.float64(metrics.iou_score(hard, tg)), dice_hard=np.float32(metrics.dice_coef(hard, tg)),
        logits3=l3.numpy(), target3=t3.numpy(), bcedice=np.float64(crit(l3, t3).item()),
        stable_bce=np.float64(losses.StableBCELoss()(l3, t3).item()),
        probs=torch.sigmoid(lg).numpy(),
    )

    # ---------------- xResidualBlock ------------------------------------------------------------
    xr = xresidualblock.xResidualBlock(16, 16, 3, 1)
    xr.load_state_dict(O.portable_state_dict(O.xresidual_block_spec(16, 16)))
    xr.train()
    xin = torch.randn(2, 16, 20, 20, generator=torch.Generator().manual_seed(8))
    yo = xr(xin)
    np.savez_compressed(os.path.join(OUT, "xresidual_2x16x20.npz"), y=yo.detach().numpy())
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print("  %-40s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


def efficientnet_golden():
    """EfficientNet encoder rows of SURVEY.md §8a: one MBConv block of each flavour, b0 extract_features, AttentiveCNN(b2)."""
    import archs  # noqa
    from efficientnet_pytorch import EfficientNet
    from efficientnet_pytorch.model import MBConvBlock
    from efficientnet_pytorch.utils import BlockArgs, get_model_params

    def grads(mod, keys):
        named = dict(mod.named_parameters())
        return {"grad:" + k: csum(named[k].grad) for k in keys}

    out = {}
    # --- single blocks (b0 global params, static padding from image_size 224) on 2 x C x 18 x 18 / 17 x 17 inputs
    _, gp = get_model_params("efficientnet-b0", None)
    cases = {"e1_k3_s1": dict(k=3, cin=32, cout=16, expand=1, stride=1, sq=8, hw=18),
             "e6_k5_s2": dict(k=5, cin=24, cout=40, expand=6, stride=2, sq=6, hw=17),
             "e6_k3_s1_skip": dict(k=3, cin=24, cout=24, expand=6, stride=1, sq=6, hw=18)}
    for name, b in cases.items():
        ba = BlockArgs(kernel_size=b["k"], num_repeat=1, input_filters=b["cin"], output_filters=b["cout"], expand_ratio=b["expand"],
                       id_skip=True, stride=1 if "skip" in name else [b["stride"]], se_ratio=0.25)   # model.py:178: repeats carry int 1
        m = MBConvBlock(ba, gp)
        sd = O.portable_state_dict(O.mbconv_spec("blk", b))
        m.load_state_dict({k[len("blk."):]: v for k, v in sd.items()})
        m.train()
        x = torch.randn(2, b["cin"], b["hw"], b["hw"], generator=torch.Generator().manual_seed(21)).requires_grad_(True)
        y = m(x)
        g = torch.randn(y.shape, generator=torch.Generator().manual_seed(22))
        (y * g).sum().backward()
        out["mb_%s:y" % name] = y.detach().numpy()
        out["mb_%s:dx" % name] = x.grad.numpy()
        for k, v in m.named_parameters():
            out["mb_%s:grad:%s" % (name, k)] = v.grad.numpy()
        out["mb_%s:bn1.running_var" % name] = m._bn1.running_var.numpy().copy()
    # --- b0 extract_features, train (drop_connect off) and eval, 2 x 3 x 64 x 64
    net = EfficientNet.from_name("efficientnet-b0", override_params={"drop_connect_rate": 0.0})
    net.load_state_dict(O.portable_state_dict(O.efficientnet_spec("efficientnet-b0")))
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(23))
    net.train()
    f = net.extract_features(x)
    f.square().mean().backward()
    out["b0:features_train"] = f.detach().numpy()
    for k in ("_conv_stem.weight", "_blocks.0._depthwise_conv.weight", "_blocks.3._se_reduce.weight", "_blocks.3._se_expand.bias",
              "_blocks.5._project_conv.weight", "_blocks.15._bn2.weight", "_conv_head.weight"):
        out["b0:csum_grad:" + k] = csum(dict(net.named_parameters())[k].grad)
    net.eval()
    with torch.no_grad():
        out["b0:features_eval"] = net.extract_features(x).numpy()
        out["b0:logits_eval"] = net(x).numpy()
    # --- AttentiveCNN on efficientnet-b2 (eval; from_name because phase_train=True needs the pretrained file)
    att = archs.AttentiveCNN({"eff_flag": True, "phase_train": False, "eff_model_name": "efficientnet-b2"})
    att.load_state_dict(O.portable_state_dict(O.attentive_cnn_spec("efficientnet-b2")))
    att.eval()
    img = torch.randn(1, 3, 96, 80, generator=torch.Generator().manual_seed(24))
    with torch.no_grad():
        out["att_b2:y_eval"] = att(img).numpy()
    np.savez_compressed(os.path.join(OUT, "efficientnet.npz"), **out)
    print("efficientnet.npz: %d bytes, %d arrays" % (os.path.getsize(os.path.join(OUT, "efficientnet.npz")), len(out)))


def tiles_golden():
    """patch_gen / patch_merge of the unmodified aerial_image_segmentation_api.py on a synthetic 150 x 200 raster."""
    import types
    for name in ("albumentations", "albumentations.augmentations", "albumentations.augmentations.transforms",
                 "albumentations.core", "albumentations.core.composition", "tensorboardX", "torchsummary"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["albumentations.core.composition"].Compose = object
    sys.modules["albumentations.core.composition"].OneOf = object
    sys.modules["albumentations.augmentations"].transforms = sys.modules["albumentations.augmentations.transforms"]
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["torchsummary"].summary = lambda *a, **k: None
    import aerial_image_segmentation_api as api
    rng = np.random.RandomState(7)
    H, W, P, C, OV = 150, 200, 64, 3, 0.5
    img = rng.randint(0, 255, size=(H, W, 3)).astype("uint8")
    patches, _ = api.patch_gen(img, img, P, OV)
    # window order: recover (h1, w1) of every patch from a coordinate image
    coord = np.stack(list(np.meshgrid(np.arange(H), np.arange(W), indexing="ij")) + [np.zeros((H, W), dtype=int)], -1)
    cp, _ = api.patch_gen(coord, coord, P, OV)
    wins = np.array([[p[0, 0, 0], p[0, 0, 1]] for p in cp], dtype=np.int32)
    # smooth probability maps (values near the 127 threshold included)
    base = rng.rand(len(patches), C, P // 8, P // 8).astype("float32")
    probs = O.tile_test_probs(base)
    merged = api.patch_merge(img, [p for p in probs], P, {"num_classes": C}, OV)
    np.savez_compressed(os.path.join(OUT, "tiles_merge_150x200.npz"), windows=wins, base=base, merged=np.stack(merged),
                        n_patches=np.int64(len(patches)), probs_csum=np.float64(probs.astype(np.float64).sum()))
    print("tiles_merge_150x200.npz: %d patches, %d bytes" % (len(patches), os.path.getsize(os.path.join(OUT, "tiles_merge_150x200.npz"))))
    # ---- the resize bridge (patch_size != map size): the reference's own patch_merge + cv2.resize on S x S maps ----
    rec = {}
    for tag, (S, P2) in {"up2": (32, 64), "up_ragged": (40, 64), "down": (96, 64), "down2": (128, 64)}.items():
        pw, _ = api.patch_gen(img, img, P2, OV)
        b = rng.rand(len(pw), C, S // 8, S // 8).astype("float32")
        pr = O.tile_test_probs(b)                                  # [P, C, S, S]
        mg = api.patch_merge(img, [p for p in pr], P2, {"num_classes": C}, OV)
        rec[tag + "_base"] = b
        rec[tag + "_merged"] = np.stack(mg)
        rec[tag + "_cfg"] = np.array([S, P2], dtype=np.int64)
    # cv2.resize itself on random uint8 rasters (the image-side shrink of get_patched_input, :361)
    import cv2
    for tag, (h, w, oh, ow) in {"r_half": (64, 64, 32, 32), "r_up": (37, 41, 50, 64), "r_down": (100, 90, 33, 47), "r_quarter": (64, 64, 16, 16)}.items():
        im = rng.randint(0, 256, size=(h, w, 3)).astype("uint8")
        rec[tag + "_src"] = im
        rec[tag + "_dst"] = cv2.resize(im, (ow, oh))
    np.savez_compressed(os.path.join(OUT, "tiles_resize_bridge.npz"), **rec)
    print("tiles_resize_bridge.npz: %d bytes" % os.path.getsize(os.path.join(OUT, "tiles_resize_bridge.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tiles":
        tiles_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "efficientnet":
        efficientnet_golden()
    else:
        main()
        efficientnet_golden()
        tiles_golden()
