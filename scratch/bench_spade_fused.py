"""Fused vs unfused SPADE forward (no_grad and grad-enabled) at the level-0 / level-1 sizes of the headline config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import normalization, ops

ssg.set_compute_dtype(torch.bfloat16)
for c, hw in ((64, 512), (128, 256)):
    mod = normalization.SPADE("spadebatch3x3", c, 3, c / 16).cuda().train()
    x = ops.to_nhwc(torch.randn(16, c, hw, hw, device="cuda"))
    for fused in (0, 1, 2):
        ops.set_spade_fused(fused)
        for grad in (False, True):
            with torch.set_grad_enabled(grad):
                xx = x.detach().requires_grad_(grad)
                for _ in range(3):
                    y = mod(xx, xx)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    y = mod(xx, xx)
                e1.record()
                torch.cuda.synchronize()
                print("C=%3d %4d^2 fused=%d grad=%d  fwd %.3f ms" % (c, hw, fused, grad, e0.elapsed_time(e1) / 10), flush=True)
    ops.set_spade_fused(False)
