"""the device-side data feed (uint8 rasters -> Normalize / Flip / layout) against the numpy restatement of
dataset.py:126-140 + albumentations' Normalize / Flip (oracle/archs_oracle.py).  Byte / elementwise work: bit-exact."""
import numpy as np
import pytest
import torch


def _batch(n, h, w, c, k, seed=0):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, size=(n, h, w, c)).astype(np.uint8)
    mask = rng.choice(np.array([0, 1, 127, 254, 255], dtype=np.uint8), size=(n, h, w, k))
    return img, mask


def test_oracle_flip_is_cv2_flip():
    """The oracle's slicing == cv2.flip for the three codes albumentations' Flip draws from."""
    cv2 = pytest.importorskip("cv2")
    import archs_oracle as A
    img, mask = _batch(3, 7, 9, 3, 3)
    codes = [0, 1, -1]
    got = A.feed_mask(mask, codes)
    for i, d in enumerate(codes):
        want = cv2.flip((mask[i].astype("float32") / 255.0).astype("uint8"), d).astype("float32").transpose(2, 0, 1)
        assert np.array_equal(got[i], want)
    assert set(np.unique(got)) <= {0.0, 1.0} and got.sum() == (mask == 255).sum()


def test_flip_code_mapping():
    from ssunet_gan_b200 import dataset
    assert [dataset.flip_code_from_cv2(d) for d in (1, 0, -1)] == [dataset.FLIP_X, dataset.FLIP_Y, dataset.FLIP_XY]


@pytest.mark.gpu
@pytest.mark.parametrize("c,mean,std", [(3, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
                                        (4, [0.485, 0.456, 0.406, 0.5], [0.229, 0.224, 0.225, 0.25]), (1, [0.5], [0.25])])
def test_feed_nchw_fp32_bit_exact(c, mean, std):
    import archs_oracle as A
    from ssunet_gan_b200 import dataset
    img, mask = _batch(5, 19, 23, c, 3, seed=c)
    cv2_codes = [None, 0, 1, -1, 1]
    codes = [dataset.FLIP_NONE if d is None else dataset.flip_code_from_cv2(d) for d in cv2_codes]
    feed = dataset.DeviceFeed(mean, std)
    x, t = feed(img, mask, flip_codes=codes, nchw=True)
    assert x.dtype == torch.float32 and tuple(x.shape) == (5, c, 19, 23) and x.is_contiguous()
    assert np.array_equal(x.cpu().numpy(), A.feed_image(img, mean, std, cv2_codes))
    assert np.array_equal(t.cpu().numpy(), A.feed_mask(mask, cv2_codes))
    assert feed.h2d_bytes == img.size + mask.size                     # bytes, not floats, cross PCIe
    # no flips, pinned host tensors as the source
    x2 = feed.images(torch.from_numpy(img).pin_memory(), nchw=True)
    assert np.array_equal(x2.cpu().numpy(), A.feed_image(img, mean, std))
    # single-class layout: mask / 1.0 (dataset.py:112,123)
    m1 = feed.masks(mask[..., :1], binarise=False)
    assert np.array_equal(m1.cpu().numpy(), mask[..., :1].astype("float32").transpose(0, 3, 1, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,impl", [(torch.bfloat16, "auto"), (torch.float32, "simt")])
def test_feed_nhwc_activation_equals_entry_conversion(dtype, impl):
    """The NHWC activation the feed writes == `ops.to_nhwc(NCHW fp32, pad_channels=True)` of the reference-layout tensor,
    bit for bit (incl. the zero padding channels), so a network fed either way returns identical logits."""
    import ssunet_gan_b200 as ssg
    import ssunet_oracle as O
    from ssunet_gan_b200 import dataset, models_seg_gan, ops
    ssg.set_compute_dtype(dtype)
    ssg.set_conv_impl(impl)
    try:
        img, _ = _batch(2, 64, 64, 3, 3, seed=9)
        feed = dataset.DeviceFeed()
        codes = [dataset.FLIP_XY, dataset.FLIP_NONE]
        a = feed.images(img, flip_codes=codes)
        ref = ops.to_nhwc(feed.images(img, flip_codes=codes, nchw=True), pad_channels=True)
        assert a.dtype == dtype and a.shape == ref.shape and a.shape[1] == ops.thin_pad(3)
        assert torch.equal(a.permute(0, 2, 3, 1).contiguous(), ref.permute(0, 2, 3, 1).contiguous())
        g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
        g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
        g = g.cuda().eval()
        with torch.no_grad():
            assert torch.equal(g(a), g(feed.images(img, flip_codes=codes, nchw=True)))
    finally:
        ssg.set_compute_dtype(torch.bfloat16)
        ssg.set_conv_impl("auto")


@pytest.mark.gpu
def test_feed_full_size_properties():
    """BASELINE-size batch (16 x 512 x 512 x 3): flipping twice is the identity; masks are {0,1} and count the 255s."""
    from ssunet_gan_b200 import dataset
    img, mask = _batch(16, 512, 512, 3, 3, seed=2)
    feed = dataset.DeviceFeed()
    d_img, d_mask = torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda()
    x = feed.images(d_img, nchw=True)
    xf = feed.images(d_img, flip_codes=[dataset.FLIP_XY] * 16, nchw=True)
    assert torch.equal(xf.flip(2, 3), x)
    t = feed.masks(d_mask)
    assert int(t.sum()) == int((mask == 255).sum()) and set(torch.unique(t).tolist()) <= {0.0, 1.0}
    tf = feed.masks(d_mask, flip_codes=[dataset.FLIP_X] * 16)
    assert torch.equal(tf.flip(3), t)
