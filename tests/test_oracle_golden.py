"""Pins oracle/ssunet_oracle.py against fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import torch

import ssunet_oracle as O

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def _csum(t):
    t = t.detach().double()
    return np.array([float(t.sum()), float(t.abs().sum()), float((t * t).sum())])


def _close(a, b, rtol=2e-4, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def test_state_dict_layout_matches_reference(golden_dir):
    lay = json.load(open(os.path.join(golden_dir, "state_layout.json")))
    g = [[k, list(s)] for k, s in O.unet_r_ss_v2_spec(3, 3, prefix="net.")]
    d = [[k, list(s)] for k, s in O.discriminator_spec(3)]
    assert g == lay["generator"]
    assert d == lay["discriminator"]
    assert len(g) == 269


def test_generator_fwd_bwd_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "generator_fwd_bwd_2x64.npz"))
    sd = O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net."))
    x, t = O.synthetic_batch(2, 3, 64, 64, seed=1234)
    _leaf = O._leafify
    _leaf(sd)
    out = O.unet_r_ss_v2(sd, x, True, prefix="net.")
    loss = O.bce_dice_loss(out, t)
    keys = O.trainable_keys(sd)
    grads = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    _close(out.detach().numpy(), z["logits"], rtol=1e-4, atol=2e-5)
    _close(float(loss.detach()), float(z["loss"]), rtol=1e-5)
    assert O.iou_score(out[:, 1:], t[:, 1:]) == float(z["iou"])
    _close(O.dice_coef(out[:, 1:], t[:, 1:]), z["dice"], rtol=1e-6)
    gmap = dict(zip(keys, grads))
    for k, c in zip(z["grad_keys"], z["grad_csum"]):
        _close(_csum(gmap[str(k)])[1:], c[1:], rtol=2e-3, atol=1e-7)
    _close(sd["net.conv0_0.bn1.running_mean"].numpy(), z["bn_running_mean"], rtol=1e-5, atol=1e-7)
    _close(sd["net.conv2_1.bn2.running_var"].numpy(), z["bn_running_var"], rtol=1e-5, atol=1e-7)


def test_generator_eval_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "generator_eval_1x96.npz"))
    sd = O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net."))
    x, _ = O.synthetic_batch(1, 3, 96, 96, seed=77)
    with torch.no_grad():
        out = O.unet_r_ss_v2(sd, x, False, prefix="net.")
    _close(out.numpy(), z["logits"], rtol=1e-4, atol=2e-6)


def test_discriminator_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "discriminator_fwd_bwd_3x96.npz"))
    sd = O.portable_state_dict(O.discriminator_spec(3))
    O._leafify(sd)
    x, _ = O.synthetic_batch(3, 3, 96, 96, seed=5)
    x.requires_grad_(True)
    lo = O.discriminator(sd, x, True)
    l = torch.nn.functional.binary_cross_entropy_with_logits(lo, torch.ones_like(lo))
    keys = O.trainable_keys(sd)
    grads = torch.autograd.grad(l, [x] + [sd[k] for k in keys])
    _close(lo.detach().numpy(), z["logit"], rtol=1e-4, atol=1e-6)
    _close(float(l), float(z["loss"]), rtol=1e-5)
    _close(grads[0].numpy(), z["dx"], rtol=1e-3, atol=1e-9)
    gmap = dict(zip(keys, grads[1:]))
    for k, c in zip(z["grad_keys"], z["grad_csum"]):
        _close(_csum(gmap[str(k)])[1:], c[1:], rtol=2e-3, atol=1e-9)


def test_gan_step_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "gan_step_2it_2x64.npz"))
    sd_g = O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net."))
    sd_d = O.portable_state_dict(O.discriminator_spec(3))
    # the oracle's generator keys carry the Generator wrapper's "net." prefix
    og = O.AdamState(O.trainable_keys(sd_g), 2e-5)
    od = O.AdamState(O.trainable_keys(sd_d), 2e-5)
    for it in range(2):
        x, t = O.synthetic_batch(2, 3, 64, 64, seed=1234 + it, blobby=(it == 1))
        r = _step_prefixed(sd_g, sd_d, og, od, x, t)
        want = z["it%d_scalars" % it]
        _close([r["loss"], r["content"], r["adv_g"], r["adv_d"]], want[:4], rtol=2e-5)
        assert r["iou"] == want[4]
        _close(r["dice"], want[5], rtol=1e-6)
        _close(r["logits"].numpy(), z["it%d_logits" % it], rtol=2e-4, atol=3e-5)
    for k, c in zip(z["g_keys"], z["g_csum"]):
        # Adam's g/sqrt(v) is sign-like on the first steps, so the plain sum carries +-lr noise
        # per near-zero-gradient element; |.| and squared sums pin the update magnitude.
        got = _csum(sd_g[str(k)].double())
        _close(got[1:], c[1:], rtol=1e-4, atol=1e-4)
        _close(got[0], c[0], rtol=1e-4, atol=4e-5 * max(1, sd_g[str(k)].numel()) ** 0.5 + 1e-4)
    for k, c in zip(z["d_keys"], z["d_csum"]):
        got = _csum(sd_d[str(k)].double())
        _close(got[1:], c[1:], rtol=1e-4, atol=1e-4)
        _close(got[0], c[0], rtol=1e-4, atol=4e-5 * max(1, sd_d[str(k)].numel()) ** 0.5 + 1e-4)
    _close(sd_g["net.final.weight"].detach().numpy(), z["final_weight"], rtol=1e-5, atol=1e-7)
    _close(sd_d["fc2.weight"].detach().numpy(), z["d_fc2_weight"], rtol=1e-5, atol=1e-7)
    assert int(sd_d["conv_blocks.1.conv_block.1.num_batches_tracked"]) == 6


def _step_prefixed(sd_g, sd_d, og, od, x, t):
    import functools
    orig = O.unet_r_ss_v2
    O.unet_r_ss_v2 = functools.partial(orig, prefix="net.")
    try:
        return O.gan_train_step(sd_g, sd_d, og, od, x, t)
    finally:
        O.unet_r_ss_v2 = orig


def test_syncbn_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "syncbn_3shards.npz"))
    x = torch.from_numpy(z["x"])
    sd = {"bn.weight": O.portable_tensor("sbn.weight", (8,)), "bn.bias": O.portable_tensor("sbn.bias", (8,)),
          "bn.running_mean": torch.zeros(8), "bn.running_var": torch.ones(8),
          "bn.num_batches_tracked": torch.zeros((), dtype=torch.int64)}
    shards = x.chunk(3, 0)

    # emulate 3 replicas: each sees a shard; the "all-reduce" sums the per-shard statistics
    def make_sync(all_shards):
        def sync(s, ss, n):
            S = sum(sh.reshape(sh.shape[0], 8, -1).sum(0).sum(-1) for sh in all_shards)
            SS = sum((sh.reshape(sh.shape[0], 8, -1) ** 2).sum(0).sum(-1) for sh in all_shards)
            return S, SS, sum(sh.shape[0] * sh[0, 0].numel() for sh in all_shards)
        return sync

    ys = []
    for sh in shards:
        sdk = dict(sd)
        ys.append(O.batch_norm(sdk, "bn", sh, True, sync_stats=make_sync(shards)))
    _close(torch.cat(ys).numpy(), z["y"], rtol=1e-5, atol=1e-6)
    _close(sdk["bn.running_mean"].numpy(), z["running_mean"], rtol=1e-6, atol=1e-8)
    _close(sdk["bn.running_var"].numpy(), z["running_var"], rtol=1e-6, atol=1e-8)
    assert int(sdk["bn.num_batches_tracked"]) == 0     # parallel path never increments it


def test_spectral_norm_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "spectral_norm_conv.npz"))
    w0 = torch.from_numpy(z["w_orig"]).requires_grad_(True)
    w1, u1, v1, sigma = O.spectral_weight(w0, torch.from_numpy(z["u0"]), torch.from_numpy(z["v0"]), True)
    _close(w1.detach().numpy(), z["w1"], rtol=1e-5, atol=1e-7)
    _close(u1.numpy(), z["u1"], rtol=1e-5, atol=1e-7)
    _close(v1.numpy(), z["v1"], rtol=1e-5, atol=1e-7)
    y = torch.nn.functional.conv2d(torch.from_numpy(z["x"]), w1, O.portable_tensor("none", (10,)) * 0, 1, 1)
    (gw,) = torch.autograd.grad(y.sum(), [w0])
    _close(gw.numpy(), z["gw_orig"], rtol=1e-3, atol=1e-5)
    assert list(z["sd_keys"]) == ["bias", "weight_orig", "weight_u", "weight_v"]


def test_metrics_and_losses_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics_loss.npz"))
    lg, tg = torch.from_numpy(z["logits"]), torch.from_numpy(z["target"])
    assert O.iou_score(lg, tg) == float(z["iou"])
    assert np.float32(O.dice_coef(lg, tg)) == z["dice"]
    assert isinstance(O.dice_coef(lg, tg), np.float32)
    hard = torch.where(lg > 0, torch.full_like(lg, 200.0), torch.full_like(lg, -200.0))
    assert O.iou_score(hard, tg) == float(z["iou_hard"])
    assert np.float32(O.dice_coef(hard, tg)) == z["dice_hard"]
    l3, t3 = torch.from_numpy(z["logits3"]), torch.from_numpy(z["target3"])
    _close(float(O.bce_dice_loss(l3, t3)), float(z["bcedice"]), rtol=1e-6)
    _close(float(O.stable_bce(l3, t3)), float(z["stable_bce"]), rtol=1e-6)
    # numpy's float32 pairwise tree, restated (what the CUDA metric kernel reproduces bit-for-bit)
    p = z["probs"].reshape(-1)
    t = z["target"].reshape(-1)
    assert O.numpy_pairwise_sum_f32(p) == p.sum()
    assert O.numpy_pairwise_sum_f32(p * t) == (p * t).sum()
    for n in (0, 1, 7, 8, 9, 127, 128, 129, 255, 1000, 4099):
        a = np.random.RandomState(n).rand(n).astype(np.float32)
        assert O.numpy_pairwise_sum_f32(a) == a.sum(), n


def test_bce_dice_nan_branch():
    x = torch.tensor([[[[float("inf"), 1.0], [0.5, -2.0]]]])
    t = torch.tensor([[[[0.0, 1.0], [1.0, 0.0]]]])
    out = O.bce_dice_loss(x, t)   # bce = inf -> 2 * dice (losses.py:297-298)
    p = torch.sigmoid(x).reshape(1, -1)
    dice = 1 - ((2 * (p * t.reshape(1, -1)).sum(1) + 1e-5) / (p.sum(1) + t.sum() + 1e-5)).sum()
    assert torch.isfinite(out)
    _close(float(out), float(2 * dice), rtol=1e-6)


def test_xresidual_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "xresidual_2x16x20.npz"))
    sd = O.portable_state_dict(O.xresidual_block_spec(16, 16))
    x = torch.randn(2, 16, 20, 20, generator=torch.Generator().manual_seed(8))
    y = O.xresidual_block(sd, x, True)
    _close(y.numpy(), z["y"], rtol=1e-4, atol=1e-5)


# ---- EfficientNet encoder rows: oracle restatement vs the unmodified reference -------------------------
MB_CASES = {"e1_k3_s1": dict(k=3, cin=32, cout=16, expand=1, stride=1, sq=8, hw=18),
            "e6_k5_s2": dict(k=5, cin=24, cout=40, expand=6, stride=2, sq=6, hw=17),
            "e6_k3_s1_skip": dict(k=3, cin=24, cout=24, expand=6, stride=1, sq=6, hw=18)}


def test_mbconv_blocks_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "efficientnet.npz"))
    for name, b in MB_CASES.items():
        sd = O.portable_state_dict(O.mbconv_spec("blk", b))
        O._leafify(sd)
        x = torch.randn(2, b["cin"], b["hw"], b["hw"], generator=torch.Generator().manual_seed(21)).requires_grad_(True)
        y = O.mbconv_block(sd, "blk", x, b, 224, True)
        g = torch.randn(y.shape, generator=torch.Generator().manual_seed(22))
        keys = [k for k in O.trainable_keys(sd)]
        grads = torch.autograd.grad((y * g).sum(), [x] + [sd[k] for k in keys])
        _close(y.detach().numpy(), z["mb_%s:y" % name], rtol=1e-4, atol=1e-5)
        _close(grads[0].numpy(), z["mb_%s:dx" % name], rtol=1e-3, atol=1e-5)
        for k, gk in zip(keys, grads[1:]):
            _close(gk.numpy(), z["mb_%s:grad:%s" % (name, k[len("blk."):])], rtol=2e-3, atol=2e-5)
        _close(sd["blk._bn1.running_var"].numpy(), z["mb_%s:bn1.running_var" % name], rtol=1e-5)


def test_efficientnet_b0_features_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "efficientnet.npz"))
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(23))
    sd = O.portable_state_dict(O.efficientnet_spec("efficientnet-b0"))
    assert len(sd) == 360                                   # reference EfficientNet-b0 state_dict size
    f = O.efficientnet_features(sd, x, "efficientnet-b0", True)
    _close(f.numpy(), z["b0:features_train"], rtol=2e-3, atol=1e-5)
    # the fixture's eval pass ran after the training pass: running statistics carry one momentum-0.01 update
    f = O.efficientnet_features(sd, x, "efficientnet-b0", False)
    _close(f.numpy(), z["b0:features_eval"], rtol=1e-3, atol=1e-5)


def test_attentive_cnn_b2_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "efficientnet.npz"))
    img = torch.randn(1, 3, 96, 80, generator=torch.Generator().manual_seed(24))
    sd = O.portable_state_dict(O.attentive_cnn_spec("efficientnet-b2"))
    y = O.attentive_cnn(sd, img, "efficientnet-b2", False)
    assert tuple(y.shape) == (1, 1024, 8, 8)          # 260 -> 130 -> 65 -> 32 -> 16 -> 8 (static pads from the nominal size)
    _close(y.numpy(), z["att_b2:y_eval"], rtol=1e-3, atol=1e-5)


def test_tile_windows_and_merge_match_reference(golden_dir):
    """patch_gen's window order and patch_merge's uint8 masks (aerial_image_segmentation_api.py:45-217)."""
    z = np.load(os.path.join(golden_dir, "tiles_merge_150x200.npz"))
    H, W, P, C, OV = 150, 200, 64, 3, 0.5
    wins = O.tile_windows(H, W, P, OV)
    assert len(wins) == int(z["n_patches"]) == 60
    assert np.array_equal(np.array(wins, dtype=np.int32), z["windows"])
    probs = O.tile_test_probs(z["base"])
    assert float(probs.astype(np.float64).sum()) == float(z["probs_csum"])
    merged = O.tile_merge(H, W, list(probs), P, C, OV)
    assert np.array_equal(np.stack(merged), z["merged"])                 # bit-identical uint8 masks
    assert set(np.unique(z["merged"])) <= {0, 255} and 0.2 < (z["merged"] == 255).mean() < 0.8


# ---------------------------------------------------------------------------------------------
# the rest of the model zoo, the supervised step, the validation step
# ---------------------------------------------------------------------------------------------
import pytest  # noqa: E402

ARCH_CASES = [("UNet", False, 32), ("NestedUNet", False, 32), ("NestedUNet", True, 32), ("SSUNet", False, 32), ("UNet_ori", False, 32),
              ("UNet_B_SS", False, 32), ("AttUNet", False, 32), ("UNet_R_SS", False, 64), ("ProgUNet", False, 32)]


def _bias_before_bn(key):
    """Conv biases that feed a BatchNorm directly (VGGBlock, conv_block, up_conv, Attention_block): zero gradient."""
    import re
    return bool(re.search(r"(^conv\d_\d\.conv[12]\.bias$)|(\.conv\.[03]\.bias$)|(\.up\.1\.bias$)|(\.(W_g|W_x|psi)\.0\.bias$)", key))


@pytest.mark.parametrize("idx", range(len(ARCH_CASES)), ids=[n + ("_ds" if d else "") for n, d, _ in ARCH_CASES])
def test_arch_zoo_oracle_matches_reference(golden_dir, idx):
    import archs_oracle as A
    name, ds, hw = ARCH_CASES[idx]
    tag = name + ("_ds" if ds else "")
    z = np.load(os.path.join(golden_dir, "archs_%s.npz" % tag))
    spec = A.arch_spec(name, ds)
    sd = O.portable_state_dict(spec, salt=idx + 1)
    O._leafify(sd)
    x, t = O.synthetic_batch(2, 3, hw, hw, seed=1234)
    out = A.arch_forward(name, sd, x, True)
    outs = out if isinstance(out, list) else [out]
    if name == "ProgUNet":
        loss = O.bce_dice_loss(outs[0], t) + sum(o.square().mean() for o in outs[1:])
    else:
        loss = A.supervised_loss(out, t)
    for i, o in enumerate(outs):
        _close(o.detach().numpy(), z["logits%d" % i], rtol=1e-4, atol=2e-5)
    _close(float(loss.detach()), float(z["loss"]), rtol=1e-5)
    keys = O.trainable_keys(sd)
    grads = dict(zip(keys, torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)))
    assert sorted(k for k in keys if grads[k] is None) == sorted(str(k) for k in z["nograd_keys"])
    for k, c in zip(z["grad_keys"], z["grad_csum"]):
        if _bias_before_bn(str(k)):       # analytically zero (BN removes the mean): both sides hold rounding noise only
            assert c[1] < 1e-4 and _csum(grads[str(k)])[1] < 1e-4
            continue
        _close(_csum(grads[str(k)])[1:], c[1:], rtol=3e-3, atol=1e-7)
    for k, c in zip(z["bn_keys"], z["bn_running_var_csum"]):
        _close(_csum(sd[str(k)]), c, rtol=1e-4, atol=1e-7)
    # eval mode + the validation loop body
    sd = O.portable_state_dict(spec, salt=idx + 1)
    xe, te = O.synthetic_batch(1, 3, hw, hw, seed=77, blobby=True)
    with torch.no_grad():
        oe = A.arch_forward(name, sd, xe, False)
    oes = oe if isinstance(oe, list) else [oe]
    for i, o in enumerate(oes):
        _close(o.numpy(), z["eval_logits%d" % i], rtol=1e-4, atol=2e-5)
    if name == "ProgUNet":
        v = (O.bce_dice_loss(oes[0], te), O.iou_score(oes[0], te), O.dice_coef(oes[0], te))
    elif isinstance(oe, list):
        v = (A.supervised_loss(oe, te), O.iou_score(oes[-1], te), O.dice_coef(oes[-1], te))
    else:
        v = (O.bce_dice_loss(oe, te), O.iou_score(oe[:, 1:3], te[:, 1:3]), O.dice_coef(oe[:, 1:3].clone(), te[:, 1:3].clone()))
    _close(float(v[0]), z["val_scalars"][0], rtol=1e-5)
    _close(float(v[1]), z["val_scalars"][1], rtol=1e-3)      # a logit within fp32 noise of 0 may flip one pixel
    _close(float(v[2]), z["val_scalars"][2], rtol=1e-5)


@pytest.mark.parametrize("name,ds,hw", [("UNet_R_SS_v2", False, 64), ("NestedUNet", True, 32)])
def test_supervised_step_oracle_matches_reference(golden_dir, name, ds, hw):
    """train.py:85-116 for two iterations: Adam(lr 1e-4, weight_decay 1e-7), weight clamp 0.7 between forward and backward."""
    import archs_oracle as A
    tag = name + ("_ds" if ds else "")
    z = np.load(os.path.join(golden_dir, "supervised_step_%s.npz" % tag))
    spec = O.unet_r_ss_v2_spec(3, 3) if name == "UNet_R_SS_v2" else A.arch_spec(name, ds)
    sd = O.portable_state_dict(spec, salt=31)
    opt = O.AdamState(O.trainable_keys(sd), lr=1e-4, weight_decay=1e-7)
    for it in range(2):
        x, t = O.synthetic_batch(2, 3, hw, hw, seed=4321 + it, blobby=(it == 1))
        r = A.supervised_train_step(name, sd, opt, x, t, clip=0.7)
        _close([r["loss"], r["dice"]], z["it%d_scalars" % it][[0, 2]], rtol=2e-5)
        _close(r["iou"], z["it%d_scalars" % it][1], rtol=1e-3)
        _close(r["logits"].numpy(), z["it%d_logits" % it], rtol=2e-4, atol=5e-5)
    for k, c in zip(z["keys"], z["csum"]):
        k = str(k)
        if sd[k].is_floating_point():
            # a conv bias in front of a BatchNorm has an analytically zero gradient; Adam turns its rounding noise into
            # +-lr steps, so its signed sum is not reproducible (and it cannot influence any output): magnitudes only
            lo = 1 if _bias_before_bn(k) else 0
            _close(_csum(sd[k])[lo:], c[lo:], rtol=2e-4, atol=1e-5)
    _close(sd[str(z["probe_key"])].detach().numpy(), z["probe"], rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# tiled inference: the cv2.resize bridge (patch size != network size)
# ---------------------------------------------------------------------------------------------
def test_cv2_resize_restatement_matches_cv2(golden_dir):
    import archs_oracle as A
    z = np.load(os.path.join(golden_dir, "tiles_resize_bridge.npz"))
    for tag in ("r_half", "r_up", "r_down", "r_quarter"):
        src, dst = z[tag + "_src"], z[tag + "_dst"]
        assert np.array_equal(A.cv2_resize_linear_u8(src, (dst.shape[1], dst.shape[0])), dst), tag
    try:
        import cv2
    except ImportError:
        return
    rng = np.random.RandomState(5)
    for (h, w, oh, ow) in ((512, 512, 1024, 1024), (33, 33, 100, 100), (16, 16, 17, 15), (5, 7, 64, 64), (128, 128, 127, 129), (64, 48, 32, 24)):
        im = rng.randint(0, 256, size=(h, w)).astype(np.uint8)
        assert np.array_equal(A.cv2_resize_linear_u8(im, (ow, oh)), cv2.resize(im, (ow, oh))), (h, w, oh, ow)


def test_tile_merge_resized_matches_reference(golden_dir):
    """oracle patch_merge with the resize bridge == the unmodified reference's patch_merge (which calls cv2.resize)."""
    import archs_oracle as A
    z = np.load(os.path.join(golden_dir, "tiles_resize_bridge.npz"))
    H, W, C, OV = 150, 200, 3, 0.5
    for tag in ("up2", "up_ragged", "down", "down2"):
        S, P2 = (int(v) for v in z[tag + "_cfg"])
        probs = O.tile_test_probs(z[tag + "_base"])
        assert probs.shape[-1] == S
        got = A.tile_merge_resized(H, W, list(probs), P2, C, OV)
        assert np.array_equal(np.stack(got), z[tag + "_merged"]), tag


# ----------------------------------------------------------------------------------------------
# headline-shape fixtures (oracle/make_golden_headline.py): 2 x 3 x 512 x 512 GAN iteration, 4-band SN7-shaped step
# ----------------------------------------------------------------------------------------------
def _sample_idx(n, sample=2048):
    step = max(1, n // sample)
    return torch.arange(0, n, step)[:sample]


def _check_grad_samples(gmap, keys, norms, samples, rtol):
    worst = 0.0
    for k, nrm, smp in zip(keys, norms, samples):
        g = gmap[str(k)].detach().reshape(-1)
        idx = _sample_idx(g.numel())
        want = torch.from_numpy(smp[:idx.numel()]).double()
        got = g[idx].double()
        if nrm < 1e-6:          # exactly-cancelling gradients (conv biases in front of a BN): rounding noise on both sides
            continue
        scale = max(float(want.norm()), nrm * (idx.numel() / g.numel()) ** 0.5)
        worst = max(worst, float((got - want).norm()) / scale)
        assert abs(float(g.double().norm()) - nrm) <= rtol * nrm, (k, float(g.double().norm()), nrm)
    assert worst < rtol, worst


def test_headline_gan_step_oracle_matches_reference(golden_dir):
    """One G+D iteration at the headline tile size (2 x 3 x 512 x 512): oracle vs the unmodified reference, logits, the six
    scalars and a strided sample + norm of every parameter gradient of both networks."""
    import functools
    z = np.load(os.path.join(golden_dir, "headline_gan_step_2x512.npz"))
    sd_g = O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net."))
    sd_d = O.portable_state_dict(O.discriminator_spec(3))
    x, t = O.synthetic_batch(2, 3, 512, 512, seed=1234, blobby=True)
    orig = O.unet_r_ss_v2
    O.unet_r_ss_v2 = functools.partial(orig, prefix="net.")
    try:
        r = O.gan_train_step(sd_g, sd_d, O.AdamState(O.trainable_keys(sd_g), 2e-5), O.AdamState(O.trainable_keys(sd_d), 2e-5), x, t)
    finally:
        O.unet_r_ss_v2 = orig
    ref = torch.from_numpy(z["logits"]).double()
    err = float((r["logits"].double() - ref).norm() / ref.norm())
    # same ATen kernels as the reference; another thread partition of the host moves the logits by `self_1thread` (2.2e-4)
    assert err < 3 * float(z["self_1thread"]) + 1e-6, err
    sc = z["scalars"]
    for k, want in zip(("loss", "content", "adv_g", "adv_d"), sc[:4]):
        _close(r[k], want, rtol=2e-4)
    _close(r["iou"], sc[4], rtol=2e-4)
    _close(r["dice"], sc[5], rtol=2e-4)
    _check_grad_samples(r["g_grads"], z["g_grad_keys"], z["g_grad_norm"], z["g_grad_sample"], 2e-2)
    _check_grad_samples(r["d_grads"], z["d_grad_keys"], z["d_grad_norm"], z["d_grad_sample"], 2e-2)
    # the yardstick the bf16 tests use: the reference's own logits move by ~sqrt(eps) under an eps input perturbation
    assert 2.5 < float(z["sens_eps_0.001"]) / float(z["sens_eps_0.0001"]) < 4.0
    assert 2.5 < float(z["sens_eps_0.0001"]) / float(z["sens_eps_1e-05"]) < 4.0


def test_sn7_four_band_step_oracle_matches_reference(golden_dir):
    """Generator(input_channels=4) forward + BCEDice + backward + clip + Adam on 2 x 4 x 64 x 64 (BASELINE configs[3] layout)."""
    z = np.load(os.path.join(golden_dir, "sn7_train_step_2x4x64.npz"))
    sd = O.portable_state_dict(O.unet_r_ss_v2_spec(3, 4, prefix="net."))
    x, t = O.synthetic_batch(2, 4, 64, 64, seed=4321, blobby=True)
    O._leafify(sd)
    out = O.unet_r_ss_v2(sd, x, True, prefix="net.")
    loss = O.bce_dice_loss(out, t)
    keys = O.trainable_keys(sd)
    grads = dict(zip(keys, torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)))
    _close(out.detach().numpy(), z["logits"], rtol=1e-4, atol=2e-5)
    _close(float(loss.detach()), float(z["loss"]), rtol=1e-5)
    _check_grad_samples(grads, z["grad_keys"], z["grad_norm"], z["grad_sample"], 5e-3)
    opt = O.AdamState(keys, 2e-5)
    opt.step(sd, grads, 0.8)
    for i, k in enumerate(z["upd_keys"]):
        _close(sd[str(k)].detach().numpy(), z["upd_%d" % i], rtol=1e-5, atol=1e-7)
