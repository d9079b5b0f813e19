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
