"""Inference leg A/B: eval-mode generator forward at batch 64 x 3 x 512^2, eval BN folded into conv1 of every BasicBlock or not."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import archs, models_seg_gan

ssg.set_compute_dtype(torch.bfloat16)
torch.manual_seed(41)
g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False}).cuda().eval()
x = torch.randn(64, 3, 512, 512, device="cuda")
for fold in (False, True, False, True):
    archs.FOLD_EVAL_BN = fold
    with torch.no_grad():
        for _ in range(2):
            y = g(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            y = g(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("fold=%s  %.2f ms per batch  %.0f img/s" % (fold, ms, 64 / ms * 1e3), flush=True)
