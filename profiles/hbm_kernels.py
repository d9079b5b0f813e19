"""Achieved HBM bandwidth of the bandwidth-bound kernels of the path (north_star: "each reported as achieved HBM GB/s
against peak").  Runs on one B200:

    python profiles/hbm_kernels.py [--out profiles/r01_hbm_kernels.json]

Every kernel is timed alone with CUDA events on the launching stream (5 warm-up + 20 timed launches, inputs larger than
the 126 MB L2 or an L2 flush between launches where they are not), and its ALGORITHMIC bytes (every live tensor read once
+ written once at its storage dtype) are divided by the mean launch time.  Peak: MEASURED_PEAKS.json
(`hbm_gbps`, burst figure for a kernel timed alone) or the B200_PROFILING.md fallback.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ssunet_gan_b200 as ssg  # noqa: E402
from ssunet_gan_b200 import _lib, ops  # noqa: E402
from ssunet_gan_b200._lib import call, dtype_code  # noqa: E402

BF = torch.bfloat16
DT = dtype_code(BF)


def timed(fn, flush, iters=20, warm=5):
    for _ in range(warm):
        fn()
    ms = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_hbm_kernels.json"))
    ap.add_argument("--only", default="", help="'bn': stop after the BatchNorm family (used by the ncu captures)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ssg.set_compute_dtype(BF)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = None
    for k in ("hbm_gbs", "hbm_gbps"):
        if k in peaks:
            peak, peak_src = float(peaks[k]), "MEASURED_PEAKS.json:" + k
            break
    if peak is None:
        peak, peak_src = 6500.0, "fallback (B200_PROFILING.md)"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    dev = "cuda"
    rows = []

    def rec(name, site, nbytes, fn, note=""):
        ms = timed(fn, flush)
        gbps = nbytes / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "reference_site": site, "algorithmic_MB": round(nbytes / 1e6, 2), "ms": round(ms, 4),
                     "GBps": round(gbps, 1), "frac_of_peak": round(gbps / peak, 3), "note": note})
        print("%-34s %9.2f MB %8.4f ms %8.1f GB/s  %.2f" % (name, nbytes / 1e6, ms, gbps, gbps / peak), flush=True)

    # ---- level-0 activation of the U-Net at the headline config: 16 x 512 x 512 x 64 bf16 = 537 MB ----
    n, h, w, c = 16, 512, 512, 64
    R = n * h * w
    x = torch.randn(R, c, device=dev).to(BF)
    y = torch.empty_like(x)
    dy = torch.randn(R, c, device=dev).to(BF)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x)
    sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
    mean = torch.zeros(c, device=dev)
    istd = torch.ones(c, device=dev)
    gam = torch.ones(c, device=dev)
    bet = torch.zeros(c, device=dev)
    E = x.numel() * 2
    rec("channel_stats (SyncBN sum/sumsq)", "batchnorm.py:59-64", E, lambda: call("ssg_channel_stats", x, DT, R, c, sums, 1))
    rec("bn_apply + ReLU", "archs.py:231", 2 * E, lambda: call("ssg_bn_apply", x, None, y, DT, R, c, mean, istd, gam, bet, 1, 0.0))
    rec("bn_apply + residual + ReLU", "archs.py:233-234", 3 * E,
        lambda: call("ssg_bn_apply", x, dy, y, DT, R, c, mean, istd, gam, bet, 1, 0.0))
    rec("bn_bwd_reduce", "autograd of batchnorm.py:73-77", 3 * E,
        lambda: call("ssg_bn_bwd_reduce", dy, y, x, DT, R, c, mean, istd, 1, 0.0, sums))
    rec("bn_bwd_apply", "autograd of batchnorm.py:73-77", 4 * E,
        lambda: call("ssg_bn_bwd_apply", dy, y, x, dx, None, DT, R, c, mean, istd, gam, sums, float(R), 1, 0.0, 1))
    rec("bn_bwd_apply + dres", "autograd of archs.py:233", 5 * E,
        lambda: call("ssg_bn_bwd_apply", dy, y, x, dx, dres, DT, R, c, mean, istd, gam, sums, float(R), 1, 0.0, 1))
    scsh = torch.cat([istd * gam, bet - mean * istd * gam]).contiguous()
    rec("bn_bwd_reduce (mask recomputed from x)", "autograd of archs.py:231 (no residual)", 2 * E,
        lambda: call("ssg_bn_bwd_reduce_rc", dy, x, DT, R, c, mean, istd, scsh, scsh[c:], 1, 0.0, sums))
    rec("bn_bwd_apply (mask recomputed from x)", "autograd of archs.py:231 (no residual)", 3 * E,
        lambda: call("ssg_bn_bwd_apply_rc", dy, x, dx, DT, R, c, mean, istd, gam, scsh, scsh[c:], sums, float(R), 1, 0.0, 1))
    if args.only == "bn":
        return
    # ---- thin -> thin 3x3 (SPADE mlp_shared at level 0): 16 x 512 x 512 x 8 bf16 both sides ----
    xt = torch.randn(R, 8, device=dev).to(BF)
    yt = torch.empty_like(xt)
    wt = torch.randn(4, 3, 3, 3, device=dev) * 0.2
    bt = torch.randn(4, device=dev) * 0.1
    dwt, dbt = torch.zeros_like(wt), torch.zeros_like(bt)
    rec("conv3x3 thin->thin fwd (3 -> 4 ch)", "normalization.py:93-95 (mlp_shared)", 2 * R * 16,
        lambda: call("ssg_conv3x3_tiny_fwd", xt, wt, bt, yt, n, h, w, 3, 4, 1, 0.0))
    rec("conv3x3 thin->thin dgrad", "autograd of normalization.py:93-95", 2 * R * 16,
        lambda: call("ssg_conv3x3_tiny_dgrad", yt, wt, xt, n, h, w, 3, 4))
    rec("conv3x3 thin->thin wgrad + bias grad", "autograd of normalization.py:93-95", 2 * R * 16,
        lambda: call("ssg_conv3x3_tiny_wgrad", xt, yt, dwt, dbt, n, h, w, 3, 4))
    del xt, yt
    gb = torch.randn(R, 2 * c, device=dev).to(BF)
    dgb = torch.empty_like(gb)
    rec("spade_modulate fwd", "normalization.py:120", 4 * E, lambda: call("ssg_spade_modulate_fwd", x, gb, y, DT, R, c))
    rec("spade_modulate bwd", "autograd of normalization.py:120", 7 * E,
        lambda: call("ssg_spade_modulate_bwd", dy, x, gb, dx, dgb, DT, R, c))
    del gb, dgb
    code = torch.empty((n, h // 2, w // 2, c), dtype=torch.uint8, device=dev)
    yp = torch.empty((n * (h // 2) * (w // 2), c), dtype=BF, device=dev)
    rec("maxpool2x2 + argmax code", "archs.py:571", E + E // 4 + E // 8, lambda: call("ssg_maxpool2x2_fwd", x, yp, code, DT, n, h, w, c))
    rec("scatter2x2 (unpool / pool bwd)", "archs.py:572", E // 4 + E // 8 + E, lambda: call("ssg_scatter2x2", yp, code, y, DT, n, h // 2, w // 2, c))
    x128 = torch.randn(n * 256 * 256, 128, device=dev).to(BF)
    up = torch.empty(n * 512 * 512, 128, dtype=BF, device=dev)
    rec("upsample2x bilinear fwd", "archs.py:573", x128.numel() * 2 * 5, lambda: call("ssg_upsample2x_fwd", x128, up, DT, n, 256, 256, 128))
    rec("upsample2x bilinear bwd", "autograd of archs.py:573", x128.numel() * 2 * 5, lambda: call("ssg_upsample2x_bwd", up, x128, DT, n, 256, 256, 128))
    cat = torch.empty(n * 512 * 512, 192, dtype=BF, device=dev)
    rec("concat2 (skip | up)", "archs.py:667", 2 * (E + up.numel() * 2), lambda: call("ssg_concat2", x, 64, up, 128, cat, DT, R))
    del x128, up, cat, yp, code
    # ---- discriminator head: adaptive 6x6 average pool of 16 x 32 x 32 x 512 + flatten, and its backward ----
    xp = torch.randn(16 * 32 * 32, 512, device=dev).to(BF)
    yp2 = torch.empty(16, 512 * 36, dtype=BF, device=dev)
    rec("adaptive_avgpool 6x6 + flatten fwd", "models_seg_gan.py:277-279", xp.numel() * 2 + yp2.numel() * 2,
        lambda: call("ssg_adaptive_avgpool_flat_fwd", xp, yp2, DT, 16, 32, 32, 512, 6, 6), "17 MB: launch-latency share")
    rec("adaptive_avgpool 6x6 + flatten bwd", "autograd of models_seg_gan.py:277-279", xp.numel() * 2 + yp2.numel() * 2,
        lambda: call("ssg_adaptive_avgpool_flat_bwd", yp2, xp, DT, 16, 32, 32, 512, 6, 6), "17 MB: launch-latency share")
    del xp, yp2
    # ---- losses / metrics on 16 x 3 x 512 x 512 fp32 logits + masks ----
    lg = torch.randn(16, 3 * 512 * 512, device=dev)
    tg = (torch.rand(16, 3 * 512 * 512, device=dev) > 0.5).float()
    ls = torch.empty(16 * 5, dtype=torch.float64, device=dev)
    out5 = torch.empty(5, device=dev)
    g3 = torch.ones(3, device=dev)
    dl = torch.empty_like(lg)
    per = 3 * 512 * 512
    rec("seg_loss_sums (BCE+Dice+MSE)", "losses.py:280-302", lg.numel() * 8, lambda: call("ssg_seg_loss_sums", lg, tg, 16, per, ls), "50 MB: fits L2, flushed")
    call("ssg_seg_loss_finalize", ls, 16, per, out5)
    rec("seg_loss_bwd", "autograd of losses.py:280-302", lg.numel() * 12, lambda: call("ssg_seg_loss_bwd", lg, tg, ls, out5, g3, 16, per, dl))
    cnt = torch.empty(2, dtype=torch.int64, device=dev)
    rec("iou_counts", "metrics.py:6-22", lg.numel() * 8, lambda: call("ssg_iou_counts", lg, tg, lg.numel(), cnt))
    # ---- clamp + Adam on the generator's 34.2 M parameters ----
    P = 34_200_000
    p_, g_, m_, v_ = (torch.randn(P, device=dev) * 0.01 for _ in range(4))
    v_.abs_()
    step = torch.zeros(1, device=dev)
    rec("clamp + Adam (G, 34.2 M params)", "srgan_utils.py:192-195 + Adam", P * 28,
        lambda: call("ssg_clamp_adam_dev", p_, g_, m_, v_, P, 2e-5, 0.9, 0.999, 1e-8, step, 0.8, 1.0))
    del p_, g_, m_, v_
    # ---- spectral norm power iteration on D block 7's weight (512 x 4608 fp32) ----
    W = torch.randn(512, 4608, device=dev) * 0.01
    u = torch.nn.functional.normalize(torch.randn(512, device=dev), dim=0)
    v = torch.nn.functional.normalize(torch.randn(4608, device=dev), dim=0)
    ws = torch.empty(512 + 4608, device=dev)
    inv = torch.empty(2, device=dev)
    rec("spectral_sigma (1 power iteration)", "spectral_norm.py:73-88", 2 * W.numel() * 4,
        lambda: call("ssg_spectral_sigma", W, u, v, 512, 4608, 1e-12, 1, inv, ws), "9.4 MB operand: latency-bound, L2 flushed")
    # ---- EfficientNet-b2 encoder, batch 16 at 260^2: stage-2 expanded activation 16 x 130 x 130 x 96 and stage-3 144 ch ----
    for (nn_, hh, cc, k, s, tag) in ((16, 130, 96, 3, 2, "b2 block1 dw3x3 s2"), (16, 65, 144, 5, 2, "b2 block3 dw5x5 s2"),
                                     (64, 65, 144, 3, 1, "b2 block2 dw3x3 s1 (batch 64)")):
        xi = torch.randn(nn_, hh, hh, cc, device=dev).to(BF)
        wdw = torch.randn(cc, 1, k, k, device=dev)
        pad = max((-(-260 // s) - 1) * s + k - 260, 0)
        oh = (hh + pad - k) // s + 1
        yo = torch.empty(nn_, oh, oh, cc, dtype=BF, device=dev)
        byt = (xi.numel() + yo.numel()) * 2
        rec("dwconv fwd   " + tag, "model.py:52-55,75", byt,
            lambda: call("ssg_dwconv2d_fwd", xi, wdw, None, yo, DT, nn_, hh, hh, cc, k, s, pad // 2, pad // 2, oh, oh))
        dxi = torch.empty_like(xi)
        rec("dwconv dgrad " + tag, "autograd", byt,
            lambda: call("ssg_dwconv2d_dgrad", yo, wdw, dxi, DT, nn_, hh, hh, cc, k, s, pad // 2, pad // 2, oh, oh))
        dwg = torch.empty_like(wdw)
        rec("dwconv wgrad " + tag, "autograd", byt,
            lambda: call("ssg_dwconv2d_wgrad", xi, yo, dwg, DT, nn_, hh, hh, cc, k, s, pad // 2, pad // 2, oh, oh))
    nn_, hh, cc, sq = 64, 65, 144, 6
    xi = torch.randn(nn_, hh, hh, cc, device=dev).to(BF)
    yo = torch.empty_like(xi)
    ps = torch.empty(nn_, cc, device=dev)
    pooled, gate = torch.empty(nn_, cc, device=dev), torch.empty(nn_, cc, device=dev)
    spre = torch.empty(nn_, sq, device=dev)
    w1, b1, w2, b2 = torch.randn(sq, cc, device=dev), torch.randn(sq, device=dev), torch.randn(cc, sq, device=dev), torch.randn(cc, device=dev)
    rec("SE plane sums (squeeze)", "model.py:80", xi.numel() * 2, lambda: call("ssg_plane_sums", xi, None, ps, DT, nn_, hh * hh, cc))
    rec("SE gate MLP", "model.py:81", (nn_ * cc * 3 + 2 * sq * cc) * 4,
        lambda: call("ssg_se_gate_fwd", ps, nn_, hh * hh, cc, sq, w1, b1, w2, b2, pooled, spre, gate), "tiny: launch-latency bound")
    rec("SE scale (excite)", "model.py:82", xi.numel() * 4, lambda: call("ssg_plane_scale", xi, gate, None, yo, DT, nn_, hh * hh, cc))
    rec("SE dgate = sum dy*x", "autograd of model.py:82", xi.numel() * 4, lambda: call("ssg_plane_sums", yo, xi, ps, DT, nn_, hh * hh, cc))
    rec("swish fwd", "utils.py:36-41", xi.numel() * 4, lambda: call("ssg_swish_fwd", xi, yo, DT, xi.numel()))
    rec("swish bwd", "utils.py:43-48", xi.numel() * 6, lambda: call("ssg_swish_bwd", yo, xi, yo, DT, xi.numel()))

    del xi, yo
    # ---- arch-zoo kernels: AttUNet / UNet_ori at 16 x 512^2, level-0 decoder tensors ----
    n, h, w, c = 16, 256, 256, 64
    xs = torch.randn(n * h * w, c, device=dev).to(BF)
    yb = torch.empty(n * 4 * h * w, c, dtype=BF, device=dev)
    rec("upsample nearest x2 fwd", "archs.py:852 (up_conv)", xs.numel() * 2 * 5, lambda: call("ssg_upsample_nearest2x_fwd", xs, yb, DT, n, h, w, c))
    rec("upsample nearest x2 bwd", "autograd of archs.py:852", xs.numel() * 2 * 5, lambda: call("ssg_upsample_nearest2x_bwd", yb, xs, DT, n, h, w, c))
    R2 = n * 512 * 512
    xg = yb                                         # 16 x 512 x 512 x 64
    zg = torch.randn(R2, device=dev).to(BF)
    yg = torch.empty_like(xg)
    dzg = torch.empty_like(zg)
    rec("attention gate fwd (x * sigmoid z)", "archs.py:142", xg.numel() * 4 + R2 * 2, lambda: call("ssg_pixel_gate_fwd", xg, zg, yg, DT, R2, c))
    rec("attention gate bwd", "autograd of archs.py:142", xg.numel() * 6 + R2 * 4,
        lambda: call("ssg_pixel_gate_bwd", yg, xg, zg, yg, dzg, DT, R2, c), "dx aliases dy (in place) for the measurement")
    del xs, yb, yg, zg, dzg
    # ---- data feed: 16 x 512 x 512 x 3 uint8 rasters + 3 mask planes ----
    img = torch.randint(0, 256, (16, 512, 512, 3), dtype=torch.uint8, device=dev)
    sub = torch.tensor([123.675, 116.28, 103.53], device=dev)
    mul = 1.0 / torch.tensor([58.395, 57.12, 57.375], device=dev)
    o8 = torch.empty(16 * 512 * 512, 8, dtype=BF, device=dev)
    o3 = torch.empty(16, 3, 512, 512, device=dev)
    rec("feed image u8 -> NHWC bf16 x8", "dataset.py:137-138 + Normalize", img.numel() + o8.numel() * 2,
        lambda: call("ssg_feed_image_u8", img, o8, DT, 0, 16, 512, 512, 3, 8, sub, mul, 0.0, None), "80 MB: fits L2, flushed")
    rec("feed image u8 -> NCHW fp32", "dataset.py:137-138 + Normalize", img.numel() + o3.numel() * 4,
        lambda: call("ssg_feed_image_u8", img, o3, 0, 1, 16, 512, 512, 3, 3, sub, mul, 0.0, None), "63 MB: fits L2, flushed")
    rec("feed mask u8 -> NCHW fp32", "dataset.py:128-140", img.numel() + o3.numel() * 4,
        lambda: call("ssg_feed_mask_u8", img, o3, 16, 512, 512, 3, None), "63 MB: fits L2, flushed")

    res = {"device": torch.cuda.get_device_name(0), "peak_GBps": peak, "peak_source": peak_src,
           "method": "CUDA events per launch, 20 timed launches after 5 warm-ups, 256 MB L2 flush between launches", "kernels": rows}
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
