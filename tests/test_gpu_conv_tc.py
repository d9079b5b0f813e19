"""tcgen05 implicit-GEMM convolution vs CPU fp32 on the same bf16-rounded operands (fwd, dgrad via autograd,
virtual concat).  Ragged tiles, multi-image tiles (NB > 1), tiny Cout (BN = 16) and 1x1 are covered."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


CASES = [
    # n, cin, cout, h, w, k
    (2, 64, 64, 16, 16, 3),
    (1, 128, 64, 24, 20, 3),      # ragged: W = 20 -> TW = 32, TH = 4
    (2, 64, 128, 8, 8, 1),        # NB = 2 images per tile
    (3, 64, 64, 2, 2, 3),         # 2x2 images, NB = 32 (mostly out of bounds)
    (1, 64, 3, 16, 16, 3),        # BN = 16, masked N store
    (1, 256, 384, 4, 4, 3),       # 3 N tiles of 128
    (1, 64, 64, 70, 130, 3),      # W > 128: two column tiles, ragged
    (1, 64, 768, 16, 16, 1),
    (2, 512, 64, 12, 12, 3),      # long K loop (72 k-blocks): pipeline wrap-around
    # stride 2 (discriminator blocks 1,3,5,7: models_seg_gan.py:267-273): element-strided TMA gather forward /
    # wgrad, parity-class data gradient
    (2, 64, 64, 16, 16, 3, 2),
    (1, 128, 128, 24, 20, 3, 2),  # ragged tiles
    (3, 64, 128, 6, 10, 3, 2),    # several images per tile
    (1, 256, 64, 34, 18, 3, 2),
    (2, 64, 64, 15, 13, 3, 2),    # odd spatial dims: last row / column handled by the parity classes
    (1, 64, 64, 140, 132, 3, 2),  # > 64 output columns: two column tiles
    # halo-form stride-2 data gradient (even sizes): more work items than SMs; one dy chunk with three dx-channel tiles (resident
    # weights reloaded mid-CTA); three dy chunks (weights streamed); ragged tiles; channel tails
    (2, 64, 64, 200, 168, 3, 2),
    (1, 192, 64, 200, 168, 3, 2),
    (1, 64, 192, 136, 104, 3, 2),
    (2, 72, 40, 36, 28, 3, 2),
    # persistent halo kernel: more work items than SMs, several N tiles per CTA (resident-weight reload when Cin <= 64)
    (1, 64, 192, 200, 168, 3),
    (2, 128, 128, 150, 100, 3),
]


@pytest.mark.parametrize("cfg", CASES)
def test_conv_tc_forward_backward(cfg):
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import ops
    n, cin, cout, h, w, k = cfg[:6]
    stride = cfg[6] if len(cfg) > 6 else 1
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")
    g = torch.Generator().manual_seed(cin * 7 + cout + h)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).bfloat16().float()
    b = torch.randn(cout, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.leaky_relu(F.conv2d(xr, wr, br, stride, k // 2), 0.2)
    gy = torch.randn(yr.shape, generator=g).bfloat16().float()
    yr.backward(gy)
    xc, wc, bc = x.cuda().requires_grad_(True), wt.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.conv2d(xc, wc, bc, stride, k // 2, ops.ACT_LEAKY, 0.2)
    y.backward(gy.cuda().bfloat16())
    assert rel(y.float(), yr) < 6e-3
    assert rel(xc.grad, xr.grad) < 8e-3       # tcgen05 dgrad (cout % 64 == 0) or SIMT dgrad
    assert rel(wc.grad, wr.grad) < 8e-3
    assert rel(bc.grad, br.grad) < 8e-3
    # cross-check against the CUDA-core kernel on the device
    ssg.set_conv_impl("simt")
    y2 = ops.conv2d(x.cuda(), wt.cuda(), b.cuda(), stride, k // 2, ops.ACT_LEAKY, 0.2)
    assert rel(y.float(), y2.float()) < 6e-3


def test_conv_tc_virtual_concat():
    from ssunet_gan_b200 import conv_tc, ops
    g = torch.Generator().manual_seed(3)
    a = torch.randn(2, 64, 20, 20, generator=g).bfloat16()
    b = torch.randn(2, 128, 20, 20, generator=g).bfloat16()
    wt = (torch.randn(64, 192, 3, 3, generator=g) / 40).bfloat16().float()
    yr = F.conv2d(torch.cat([a.float(), b.float()], 1), wt, None, 1, 1)
    ac, bc = ops.to_nhwc(a.cuda().float()), ops.to_nhwc(b.cuda().float())
    y = ops.empty_nhwc(2, 64, 20, 20, torch.bfloat16)
    conv_tc.forward(ac, wt.cuda(), None, y, 1, 1, 0, 0.0, x1=bc)
    assert rel(y.float(), yr) < 6e-3


@pytest.mark.parametrize("cfg", [(2, 64, 64, 64, 48), (1, 64, 128, 200, 168), (1, 128, 128, 40, 24)])
def test_conv_tc_stride2_dgrad_carries_producer_activation(cfg):
    """models_seg_gan.py:38-39,52-57: conv -> LeakyReLU -> stride-2 conv.  The stride-2 data gradient multiplies by the LeakyReLU
    derivative of the first layer's output in its epilogue (`input_act`); compare with autograd on the unfused fp32 chain."""
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import ops
    n, c1, c2, h, w = cfg
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")
    g = torch.Generator().manual_seed(c1 + c2 + h)
    x = torch.randn(n, 16, h, w, generator=g).bfloat16().float()
    w1 = (torch.randn(c1, 16, 3, 3, generator=g) / 12.0).bfloat16().float()
    w2 = (torch.randn(c2, c1, 3, 3, generator=g) / math.sqrt(9 * c1)).bfloat16().float()
    xr, w1r, w2r = x.clone().requires_grad_(True), w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    yr = F.conv2d(F.leaky_relu(F.conv2d(xr, w1r, None, 1, 1), 0.2), w2r, None, 2, 1)
    gy = torch.randn(yr.shape, generator=g).bfloat16().float()
    yr.backward(gy)
    xc, w1c, w2c = x.cuda().requires_grad_(True), w1.cuda().requires_grad_(True), w2.cuda().requires_grad_(True)
    mid = ops.conv2d(xc, w1c, None, 1, 1, ops.ACT_LEAKY, 0.2)
    y = ops.conv2d(mid, w2c, None, 2, 1, ops.ACT_NONE, 0.0, input_act=(ops.ACT_LEAKY, 0.2))
    y.backward(gy.cuda().bfloat16())
    assert rel(y.float(), yr) < 8e-3
    assert rel(xc.grad, xr.grad) < 1.2e-2
    assert rel(w1c.grad, w1r.grad) < 1.2e-2
    assert rel(w2c.grad, w2r.grad) < 8e-3


THIN = [
    # n, cin, cout, h, w, k, stride   -- thin (channel-padded) operands: image stems, SPADE maps, logits head
    (2, 3, 64, 20, 24, 3, 1),     # archs.py:576 conv0_0.conv1 / models_seg_gan.py:267 block 0
    (2, 3, 64, 20, 24, 1, 1),     # BasicBlock shortcut on the image
    (1, 64, 3, 16, 16, 3, 1),     # SPADE.x2map (normalization.py:94)
    (2, 3, 4, 12, 12, 3, 1),      # SPADE.mlp_shared, both sides thin: csrc/conv_tiny.cu (one pixel per thread)
    (2, 3, 8, 30, 22, 3, 1),      # level 1 (nhidden = 8)
    (1, 4, 3, 17, 9, 3, 1),       # roles swapped
    (1, 8, 8, 64, 40, 3, 1),
    (3, 5, 7, 9, 11, 3, 1),       # odd channel counts inside the 8-channel storage
    (2, 3, 4, 100, 70, 3, 1),     # width not a multiple of the 4- / 8-pixel runs
    (1, 4, 128, 18, 10, 3, 1),    # SPADE gamma|beta at level 0
    (1, 24, 768, 8, 8, 3, 1),     # level 3
    (2, 48, 192, 6, 6, 3, 1),
    (1, 64, 3, 33, 17, 1, 1),     # final 1x1 head (archs.py:615)
    (2, 3, 64, 16, 16, 3, 2),
    (2, 4, 128, 160, 160, 3, 1),  # SPADE gamma|beta, 400 work items over 148 persistent CTAs: weights reloaded mid-CTA
]


@pytest.mark.parametrize("cfg", THIN)
def test_conv_tc_thin_channels(cfg):
    """Thin tensors are stored with their channel count rounded up to 8 (zeros) and run through the same tcgen05 kernels;
    results and all three gradients must match the unpadded convolution."""
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import ops
    n, cin, cout, h, w, k, stride = cfg
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")
    g = torch.Generator().manual_seed(cin * 13 + cout + h)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).bfloat16().float()
    b = torch.randn(cout, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wr, br, stride, k // 2))
    gy = torch.randn(yr.shape, generator=g).bfloat16().float()
    yr.backward(gy)
    xc, wc, bc = x.cuda().requires_grad_(True), wt.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    xs = ops.to_nhwc(xc, pad_channels=True)
    assert xs.shape[1] == ops.thin_pad(cin) and xs.shape[1] % 8 == 0
    ys = ops.conv2d(xs, wc, bc, stride, k // 2, ops.ACT_RELU, 0.0, cout_store=ops.thin_pad(cout))
    assert ys.shape[1] == ops.thin_pad(cout)
    if ys.shape[1] > cout:
        assert float(ys[:, cout:].float().abs().max()) == 0.0          # padding channels stay zero
    y = ops.to_nchw_f32(ys, channels=cout)
    y.backward(gy.cuda())
    assert rel(y, yr) < 6e-3
    assert rel(xc.grad, xr.grad) < 8e-3
    assert rel(wc.grad, wr.grad) < 8e-3
    assert rel(bc.grad, br.grad) < 8e-3


@pytest.mark.parametrize("cfg", [(2, 64, 64, 40, 24, 3), (1, 128, 192, 19, 13, 3), (3, 64, 128, 8, 8, 1), (1, 192, 64, 33, 9, 3)])
def test_conv_epilogue_bn_statistics(cfg):
    """The per-channel sum / sum-of-squares the conv epilogue emits (batchnorm.py:59-64 as a by-product) equal the
    standalone reduction kernel's result on the stored bf16 output, including ragged tiles (masked rows)."""
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import ops
    from ssunet_gan_b200._lib import call, dtype_code
    n, cin, cout, h, w, k = cfg
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")
    g = torch.Generator().manual_seed(7 * cin + cout)
    x = torch.randn(n, cin, h, w, generator=g).cuda()
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).cuda()
    y, sums = ops.conv2d(x, wt, None, 1, k // 2, want_stats=True)
    assert sums is not None and sums.dtype == torch.float64 and sums.numel() == 2 * cout
    ref = torch.empty(2 * cout, dtype=torch.float64, device="cuda")
    call("ssg_channel_stats", y, dtype_code(y.dtype), n * h * w, cout, ref, 1)
    yf = y.float()
    want = torch.cat([yf.sum((0, 2, 3)), (yf * yf).sum((0, 2, 3))]).double()
    assert rel(ref, want) < 1e-5
    assert rel(sums, want) < 1e-5
    assert float((sums[cout:] - want[cout:]).abs().max() / want[cout:].abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------------------------------
# data gradient accumulated into an existing buffer (TMA reduce-add epilogue) and the module-level GradSink
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(2, 64, 64, 16, 16, 3), (1, 192, 64, 40, 24, 1), (1, 128, 256, 19, 13, 3), (2, 64, 8, 33, 20, 3),
                                 (1, 8, 128, 24, 24, 3), (1, 64, 192, 200, 168, 3)])
def test_conv_tc_dgrad_accumulate(cfg):
    """dx += dgrad(dy) inside the kernel == bf16(dx + bf16(dgrad(dy))): the second consumer's contribution lands in the
    buffer the first one wrote (cin / cout here are the STORED channel counts, thin ones included)."""
    from ssunet_gan_b200 import conv_tc, ops
    n, cin, cout, h, w, k = cfg
    pad = k // 2
    assert conv_tc.can_accumulate(k, 1, pad) and not conv_tc.can_accumulate(3, 2, 1)
    g = torch.Generator().manual_seed(sum(cfg))
    wgt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).cuda()
    dy = ops.to_nhwc(torch.randn(n, cout, h, w, generator=g).cuda(), torch.bfloat16)
    base = ops.to_nhwc(torch.randn(n, cin, h, w, generator=g).cuda(), torch.bfloat16)
    plain = ops.empty_nhwc(n, cin, h, w, torch.bfloat16)
    conv_tc.dgrad(dy, wgt, plain, 1, pad)
    acc = base.clone(memory_format=torch.preserve_format)
    assert ops.is_nhwc(acc)
    conv_tc.dgrad(dy, wgt, acc, 1, pad, accumulate=True)
    want = (base.float() + plain.float()).to(torch.bfloat16)
    torch.cuda.synchronize()
    diff = (acc.float() - want.float()).abs()
    # identical up to the rounding mode of the L2's bf16 adder: at most one bf16 ulp of the result
    assert bool((diff <= want.float().abs() * 2.0 ** -7 + 1e-30).all()), float(diff.max())
    assert rel(acc.float(), want.float()) < 2e-3


@pytest.mark.parametrize("which", ["basic_block", "spade"])
def test_grad_sink_equals_autograd_accumulation(which, monkeypatch):
    """BasicBlock (conv1 + shortcut) and SPADE (x2map + modulation) with the shared gradient buffer against the same module
    with the sink switched off (autograd adds the two contributions): same input gradient, same parameter gradients."""
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import archs, normalization, ops
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")
    torch.manual_seed(7)
    if which == "basic_block":
        mod = archs.BasicBlock(192, 64).cuda().train()
        x0 = torch.randn(2, 192, 40, 24)
        run = lambda t: mod(t)
    else:
        mod = normalization.SPADE("spadebatch3x3", 128, 3, 128 / 16).cuda().train()
        x0 = torch.randn(2, 128, 40, 24)
        run = lambda t: mod(t, t)
    gy = torch.randn(2, 64 if which == "basic_block" else 128, 40, 24)

    def once():
        x = ops.to_nhwc(x0.cuda()).detach().requires_grad_(True)
        xin = ops.relu(x)                        # a non-leaf producer, so the module's input gradient flows through autograd
        for p in mod.parameters():
            p.grad = None
        y = run(xin)
        y.backward(ops.to_nhwc(gy.cuda()))
        return y.detach().float(), x.grad.detach().float().clone(), {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}

    made = []
    orig = ops.grad_sink_for
    monkeypatch.setattr(ops, "grad_sink_for", lambda t, expected=2: made.append(orig(t, expected)) or made[-1])
    y1, dx1, g1 = once()
    assert len(made) == 1 and made[0] is not None and made[0].buf is None and made[0].arrived == 0     # used and reset
    monkeypatch.setattr(ops, "grad_sink_for", lambda t, expected=2: None)
    y2, dx2, g2 = once()
    assert torch.equal(y1, y2)
    assert rel(dx1, dx2) < 2e-3, rel(dx1, dx2)
    assert g1.keys() == g2.keys()
    for k in g1:
        assert rel(g1[k], g2[k]) < 1e-5, (k, rel(g1[k], g2[k]))


@pytest.mark.parametrize("cfg", [(2, 64, 128, 64, 16, 16, 3), (1, 128, 256, 128, 40, 24, 1), (1, 384, 384, 384, 19, 13, 3), (1, 64, 8, 64, 33, 20, 3)])
def test_conv_tc_dgrad_split_outputs(cfg):
    """Data gradient written straight into the two sources of a virtual concatenation == slices of the plain data gradient."""
    from ssunet_gan_b200 import conv_tc, ops
    n, c0, c1, cout, h, w, k = cfg
    pad = k // 2
    g = torch.Generator().manual_seed(sum(cfg))
    wgt = (torch.randn(cout, c0 + c1, k, k, generator=g) / math.sqrt((c0 + c1) * k * k)).cuda()
    dy = ops.to_nhwc(torch.randn(n, cout, h, w, generator=g).cuda(), torch.bfloat16)
    full = ops.empty_nhwc(n, c0 + c1, h, w, torch.bfloat16)
    conv_tc.dgrad(dy, wgt, full, 1, pad)
    d0, d1 = ops.empty_nhwc(n, c0, h, w, torch.bfloat16), ops.empty_nhwc(n, c1, h, w, torch.bfloat16)
    conv_tc.dgrad_split(dy, wgt, d0, d1, 1, pad)
    torch.cuda.synchronize()
    assert torch.equal(d0, full[:, :c0]) and torch.equal(d1, full[:, c0:])
    conv_tc.dgrad_split(dy, wgt, d0, d1, 1, pad, accumulate=True)           # d += d: exactly 2 x in bf16
    torch.cuda.synchronize()
    assert torch.equal(d0.float(), 2 * full[:, :c0].float()) and torch.equal(d1.float(), 2 * full[:, c0:].float())


def test_basic_block_virtual_concat_equals_materialised():
    """Decoder BasicBlock on CatPair(skip, up) (no torch.cat, no split in the backward) against the same block on the
    materialised concatenation: outputs, both input gradients and every parameter gradient."""
    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import archs, ops
    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl("auto")
    torch.manual_seed(11)
    mod = archs.BasicBlock(64 + 128, 64).cuda().train()
    a0, b0 = torch.randn(2, 64, 40, 24), torch.randn(2, 128, 40, 24)
    gy = torch.randn(2, 64, 40, 24)

    def once(virtual):
        a = ops.to_nhwc(a0.cuda()).detach().requires_grad_(True)
        b = ops.to_nhwc(b0.cuda()).detach().requires_grad_(True)
        for p in mod.parameters():
            p.grad = None
        x = ops.concat_channels(ops.relu(a), ops.relu(b), virtual=virtual)
        assert isinstance(x, ops.CatPair) == virtual
        l0 = ops._lib.launch_count
        y = mod(x)
        y.backward(ops.to_nhwc(gy.cuda()))
        return (y.detach().float(), a.grad.float().clone(), b.grad.float().clone(),
                {k: p.grad.clone() for k, p in mod.named_parameters()}, ops._lib.launch_count - l0)

    y1, da1, db1, g1, n1 = once(True)
    y2, da2, db2, g2, n2 = once(False)
    assert rel(y1, y2) < 1e-3
    assert rel(da1, da2) < 3e-3 and rel(db1, db2) < 3e-3, (rel(da1, da2), rel(db1, db2))
    for k in g1:
        assert rel(g1[k], g2[k]) < 2e-3, (k, rel(g1[k], g2[k]))
    with torch.no_grad():      # inference path
        mod.eval()
        ye = mod(ops.concat_channels(ops.to_nhwc(a0.cuda()), ops.to_nhwc(b0.cuda()), virtual=True))
        yr = mod(ops.concat_channels(ops.to_nhwc(a0.cuda()), ops.to_nhwc(b0.cuda())))
        assert rel(ye.float(), yr.float()) < 1e-3
