"""Supervised-step throughput of every network in `archs.__all__` (SURVEY §8f rows 2 + 4) on one B200:

    python profiles/arch_zoo_bench.py [--batch 8] [--size 512] [--out profiles/r01_arch_zoo.json]

One step = train.py:85-116 (`train_step.supervised_train_step`): forward, BCEDiceLoss, weight clamp, backward,
Adam(lr 1e-4, weight_decay 1e-7) -- bf16 tensor-core path, synthetic tiles resident on the device, metrics off.
Timing: 3 warm-up + 5 timed steps, CUDA events on the launching stream.  The convolution FLOPs (2 * MACs of every
tcgen05 conv launch, counted by the host binding) of one step are divided by the step time: whole-step tensor rate.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ssunet_gan_b200 as ssg  # noqa: E402
from ssunet_gan_b200 import _lib, archs, losses, train_step  # noqa: E402

CONV = ("ssg_conv2d_fwd_tc", "ssg_conv2d_dgrad_tc", "ssg_conv2d_dgrad_tc_acc", "ssg_conv2d_dgrad_tc_split", "ssg_conv2d_wgrad_tc", "ssg_conv2d_wgrad_tc_acc")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_arch_zoo.json"))
    a = ap.parse_args()
    torch.cuda.set_device(0)
    ssg.set_compute_dtype(torch.bfloat16)
    cfg = {"num_classes": 3, "deep_supervision": False, "optimizer": "Adam", "lr": 1e-4, "weight_decay": 1e-7, "clip": 0.7}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(a.batch, 3, a.size, a.size, generator=g).cuda()
    t = (torch.rand(a.batch, 3, a.size, a.size, generator=g) > 0.5).float().cuda()
    crit = losses.BCEDiceLoss()
    rows = []
    for name in archs.__all__:
        torch.manual_seed(41)
        net = archs.__dict__[name](3, 3, False).cuda().train()
        opt = train_step.make_supervised_optimizer(net, cfg)
        for _ in range(3):
            train_step.supervised_train_step(cfg, net, crit, opt, x, t, with_metrics=False)
        _lib.profile_reset(CONV)
        l0 = _lib.launch_count
        train_step.supervised_train_step(cfg, net, crit, opt, x, t, with_metrics=False)
        launches = _lib.launch_count - l0
        prof = _lib.profile_collect()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            r = train_step.supervised_train_step(cfg, net, crit, opt, x, t, with_metrics=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        row = {"arch": name, "params_M": round(sum(p.numel() for p in net.parameters()) / 1e6, 2), "batch": a.batch, "size": a.size,
               "ms_per_step": round(ms, 2), "img_per_s": round(a.batch / ms * 1e3, 1), "loss": round(float(r["loss"]), 5),
               "conv_gflop_per_step": round(prof["flops"] / 1e9, 1), "conv_kernel_ms": round(prof["ms"], 2),
               "conv_tflops_in_kernels": round(prof["flops"] / (prof["ms"] * 1e-3) / 1e12, 1) if prof["ms"] else None,
               "step_tflops": round(prof["flops"] / (ms * 1e-3) / 1e12, 1), "abi_calls_per_step": launches,
               "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del net, opt, r
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    with open(a.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(0), "dtype": "bf16", "launch": "eager (one C-ABI call per kernel)",
                   "step": "train.py:85-116 via train_step.supervised_train_step, metrics off", "rows": rows}, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
