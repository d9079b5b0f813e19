"""Tiled inference (reference aerial_image_segmentation_api.py:30-217, 376-404): device-side patch merge
against the fixture written by the unmodified reference and against the numpy oracle -- uint8 masks, bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_patch_gen_matches_reference_windows(golden_dir):
    import ssunet_oracle as O
    from ssunet_gan_b200 import aerial_image_segmentation_api as api
    z = np.load(os.path.join(golden_dir, "tiles_merge_150x200.npz"))
    H, W, P = 150, 200, 64
    img = np.arange(H * W * 3, dtype=np.int64).reshape(H, W, 3)
    patches, masks = api.patch_gen(img, img, P, 0.5)
    assert len(patches) == int(z["n_patches"])
    got = np.array([[p[0, 0, 0] // (W * 3), (p[0, 0, 0] // 3) % W] for p in patches], dtype=np.int32)
    assert np.array_equal(got, z["windows"])
    assert all(p.shape == (P, P, 3) for p in patches) and np.array_equal(masks[7], patches[7])


def test_patch_merge_bit_exact_vs_reference(golden_dir):
    import ssunet_oracle as O
    from ssunet_gan_b200 import aerial_image_segmentation_api as api
    z = np.load(os.path.join(golden_dir, "tiles_merge_150x200.npz"))
    H, W, P, C, OV = 150, 200, 64, 3, 0.5
    probs = O.tile_test_probs(z["base"])
    img = np.zeros((H, W, 3), dtype=np.uint8)
    merged = api.patch_merge(img, [p for p in probs], P, {"num_classes": C}, OV)
    assert len(merged) == C and merged[0].dtype == np.uint8 and merged[0].shape == (H, W)
    assert np.array_equal(np.stack(merged), z["merged"])
    # same through CUDA tensors, one batch
    merged2 = api.patch_merge(img, torch.from_numpy(probs).cuda(), P, {"num_classes": C}, OV)
    assert np.array_equal(np.stack(merged2), z["merged"])


def test_patch_merge_ragged_sizes_vs_oracle():
    """Rasters that are not a multiple of the step (windows overlap unevenly; every pixel still covered) and a raster of
    exactly one patch, against the numpy restatement."""
    import ssunet_oracle as O
    from ssunet_gan_b200 import aerial_image_segmentation_api as api
    rng = np.random.RandomState(3)
    for (H, W, P, C, OV) in ((97, 131, 32, 2, 0.5), (64, 64, 64, 3, 0.5), (120, 88, 40, 1, 0.25)):
        wins = O.tile_windows(H, W, P, OV)
        base = rng.rand(len(wins), C, P // 8, P // 8).astype(np.float32)
        probs = O.tile_test_probs(base)
        want = O.tile_merge(H, W, list(probs), P, C, OV)
        got = api.patch_merge(np.zeros((H, W, 3), np.uint8), [p for p in probs], P, {"num_classes": C}, OV)
        assert np.array_equal(np.stack(got), np.stack(want)), (H, W, P)


def test_segmentation_inference_batched_matches_per_patch_oracle_merge():
    """End to end: batched forward of the generator over all windows + device merge == the reference's per-patch loop
    (sigmoid on the device, host merge) applied to the SAME logits."""
    import ssunet_oracle as O
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import aerial_image_segmentation_api as api, models_seg_gan
    ssg.set_compute_dtype(torch.bfloat16)
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
    g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
    g.cuda().eval()
    H, W, P, OV = 160, 224, 64, 0.5
    rng = np.random.RandomState(11)
    raster = rng.randn(H, W, 3).astype(np.float32)
    patches, _ = api.patch_gen(raster, raster, P, OV)
    patch_set = np.stack([p.transpose(2, 0, 1) for p in patches]).astype(np.float32)
    config = {"patch_size": P, "input_w": P, "patch_overlap": OV, "num_classes": 3}
    got, got_gt = api.segmentation_inference(g, raster, patch_set, patch_set, config, False, batch_size=16)
    # reference flow on the same logits: batch-1 semantics are the same computation per patch in eval mode
    with torch.no_grad():
        logits = torch.cat([g(torch.from_numpy(patch_set[i:i + 16]).cuda()) for i in range(0, len(patch_set), 16)])
        probs = torch.sigmoid(logits).cpu().numpy()
    want = O.tile_merge(H, W, list(probs), P, 3, OV)
    diff = sum(int((a != b).sum()) for a, b in zip(got, want))
    # torch.sigmoid vs the kernel's 1 / (1 + expf(-x)) may differ in the last ulp: only a probability within one ulp of
    # 128 / 255 could flip a vote
    assert diff <= 2, diff
    assert got_gt is got or all(np.array_equal(a, b) for a, b in zip(got, got_gt))


# ------------------------------------------------------------------------------------------------------------------
# the cv2.resize bridge (patch size != network size, e.g. config_v1.json: 1024 vs 512)
# ------------------------------------------------------------------------------------------------------------------
def test_resize_u8_bit_exact_vs_cv2_fixtures(golden_dir):
    from ssunet_gan_b200 import aerial_image_segmentation_api as api
    z = np.load(os.path.join(golden_dir, "tiles_resize_bridge.npz"))
    for tag in ("r_half", "r_up", "r_down", "r_quarter"):
        src, want = z[tag + "_src"], z[tag + "_dst"]
        batch = torch.from_numpy(np.stack([src, src[::-1].copy()])).cuda()
        got = api.resize_u8(batch, want.shape[0], want.shape[1]).cpu().numpy()
        assert got.dtype == np.uint8 and np.array_equal(got[0], want), tag
    import archs_oracle as A            # a larger raster against the numpy restatement (itself pinned to cv2)
    rng = np.random.RandomState(8)
    big = rng.randint(0, 256, size=(3, 256, 256, 3)).astype(np.uint8)
    for (oh, ow) in ((128, 128), (512, 512), (200, 300)):
        got = api.resize_u8(torch.from_numpy(big).cuda(), oh, ow).cpu().numpy()
        for i in range(3):
            assert np.array_equal(got[i], A.cv2_resize_linear_u8(big[i], (ow, oh))), (oh, ow, i)


def test_patch_merge_with_resize_bridge_bit_exact_vs_reference(golden_dir):
    """Maps smaller / larger than the patch: the device merge == the unmodified reference's patch_merge (cv2.resize inside)."""
    import ssunet_oracle as O
    from ssunet_gan_b200 import aerial_image_segmentation_api as api
    z = np.load(os.path.join(golden_dir, "tiles_resize_bridge.npz"))
    H, W, C, OV = 150, 200, 3, 0.5
    img = np.zeros((H, W, 3), dtype=np.uint8)
    for tag in ("up2", "up_ragged", "down", "down2"):
        S, P2 = (int(v) for v in z[tag + "_cfg"])
        probs = O.tile_test_probs(z[tag + "_base"])
        merged = api.patch_merge(img, torch.from_numpy(probs).cuda(), P2, {"num_classes": C}, OV)
        assert np.array_equal(np.stack(merged), z[tag + "_merged"]), tag


def test_get_patched_input_device_vs_oracle():
    """Image side of the bridge: windows, cv2.resize shrink, Normalize(), / 255, CHW -- bit-exact against the numpy restatement."""
    import archs_oracle as A
    from ssunet_gan_b200 import aerial_image_segmentation_api as api
    rng = np.random.RandomState(4)
    raster = rng.randint(0, 256, size=(150, 200, 3)).astype(np.uint8)
    for (P, S) in ((64, 32), (64, 48), (64, 64)):
        cfg = {"patch_size": P, "input_w": S, "input_h": S, "patch_overlap": 0.5, "num_classes": 3}
        img, patches, raw = api.get_patched_input(raster, cfg, False, chunk=7)
        want = A.patched_input(raster, P, S, 0.5)
        assert img is raster and patches.dtype == torch.float32 and tuple(patches.shape) == want.shape
        assert np.array_equal(patches.cpu().numpy(), want), (P, S)
        assert raw.shape == (want.shape[0], P, P, 3)


def test_segmentation_inference_with_resize_bridge():
    """End to end at patch 128 / network 64: device shrink + batched forward + resizing vote == the reference flow (cv2-style
    resize of the uint8 sigmoid maps, host merge) applied to the SAME logits."""
    import archs_oracle as A
    import ssunet_oracle as O
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import aerial_image_segmentation_api as api, models_seg_gan
    ssg.set_compute_dtype(torch.bfloat16)
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
    g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
    g.cuda().eval()
    H, W, P, S, OV = 300, 420, 128, 64, 0.5
    raster = np.random.RandomState(12).randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    cfg = {"patch_size": P, "input_w": S, "input_h": S, "patch_overlap": OV, "num_classes": 3}
    img, patch_set, raw = api.get_patched_input(raster, cfg, False)
    got, _ = api.segmentation_inference(g, img, patch_set, raw, cfg, False, batch_size=16)
    with torch.no_grad():
        logits = torch.cat([g(patch_set[i:i + 16]) for i in range(0, len(patch_set), 16)])
        probs = torch.sigmoid(logits).cpu().numpy()
    want = A.tile_merge_resized(H, W, list(probs), P, 3, OV)
    assert got[0].shape == (H, W) and got[0].dtype == np.uint8
    diff = sum(int((a != b).sum()) for a, b in zip(got, want))
    # sigmoid in the kernel vs torch.sigmoid may differ in the last ulp: a probability within one ulp of a uint8 step can move
    # one quantised sample by 1, which only matters where the interpolated value sits exactly at the 127 threshold
    assert diff <= 8, diff
