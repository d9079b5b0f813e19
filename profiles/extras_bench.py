"""Throughput of the two path components north_star names that are NOT in the headline step's default wiring (one B200):

    python profiles/extras_bench.py [--out profiles/r02_extras.json]

  * the G+D step with the SPECTRAL-NORM discriminator (`spectral_norm.apply_spectral_norm_to_discriminator`: one power
    iteration + W / sigma per wrapped layer and per forward, spectral_norm.py:38-88; three D forwards per step), same batch
    16 x 3 x 512 x 512 as the headline, eager launches (the power-iteration buffers u / v change every forward);
  * the EfficientNet-b2 encoder as `archs.AttentiveCNN` wires it (archs.py:409-466: bilinear resize to 260 x 260 ->
    `extract_features` -> 1x1 `conv_a` to 1024 channels): forward + backward in train mode and forward in eval mode, batch 16
    of 3 x 512 x 512 tiles.
Timing: 3 warm-up + 5 timed iterations, CUDA events on the launching stream, inputs resident on the device.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ssunet_gan_b200 as ssg  # noqa: E402
from ssunet_gan_b200 import _lib, archs, models_seg_gan, optim, spectral_norm, train_step  # noqa: E402


def timed(fn, steps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_extras.json"))
    a = ap.parse_args()
    torch.cuda.set_device(0)
    ssg.set_compute_dtype(torch.bfloat16)
    gen = torch.Generator().manual_seed(1234)
    x = torch.randn(a.batch, 3, a.size, a.size, generator=gen).cuda()
    t = (torch.rand(a.batch, 3, a.size, a.size, generator=gen) > 0.5).float().cuda()
    out = {"device": torch.cuda.get_device_name(0), "batch": a.batch, "size": a.size, "dtype": "bf16"}

    # ---- G+D step, plain vs spectral-norm discriminator (eager launches in both cases, for a like-for-like comparison)
    for tag, sn in (("gan_step_plain_discriminator", False), ("gan_step_spectral_norm_discriminator", True)):
        torch.manual_seed(41)
        g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
        d = models_seg_gan.Discriminator(3)
        if sn:
            spectral_norm.apply_spectral_norm_to_discriminator(d)
        g.cuda().train(); d.cuda().train()
        og = optim.FusedClampAdam(g.parameters(), lr=2e-5)
        od = optim.FusedClampAdam(d.parameters(), lr=2e-5)
        l0 = _lib.launch_count
        ms = timed(lambda: train_step.gan_train_step(g, d, og, od, x, t, with_metrics="device"))
        calls = (_lib.launch_count - l0) // 8
        out[tag] = {"img_per_s": round(a.batch / (ms / 1e3), 1), "ms_per_step": round(ms, 3), "c_abi_calls_per_step": calls,
                    "launch": "eager"}
        print(tag, out[tag], flush=True)
        del g, d, og, od
        torch.cuda.empty_cache()

    # ---- EfficientNet-b2 encoder (AttentiveCNN)
    torch.manual_seed(41)
    enc = archs.AttentiveCNN({"eff_flag": True, "eff_model_name": "efficientnet-b2", "phase_train": False}).cuda()
    opt = optim.FusedClampAdam(enc.parameters(), lr=1e-4)

    def fwd_bwd():
        opt.zero_grad()
        y = enc(x)
        y.square().mean().backward()

    enc.train()
    ms = timed(fwd_bwd)
    out["efficientnet_b2_encoder_train_fwd_bwd"] = {"img_per_s": round(a.batch / (ms / 1e3), 1), "ms": round(ms, 3),
                                                    "what": "AttentiveCNN(efficientnet-b2): resize 512 -> 260, extract_features, conv_a; fwd + bwd"}
    print(out["efficientnet_b2_encoder_train_fwd_bwd"], flush=True)
    enc.eval()

    def fwd():
        with torch.no_grad():
            enc(x)

    ms = timed(fwd)
    out["efficientnet_b2_encoder_eval_fwd"] = {"img_per_s": round(a.batch / (ms / 1e3), 1), "ms": round(ms, 3)}
    print(out["efficientnet_b2_encoder_eval_fwd"], flush=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
