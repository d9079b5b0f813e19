"""Micro-benchmark of the tcgen05 conv kernels on generator / discriminator shapes (CUDA events, L2 flushed by size)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, conv_tc

B = int(os.environ.get("B", "16"))
SHAPES = [  # name, cin, cout, hw, k, stride
    ("conv0_0.conv2", 64, 64, 512, 3, 1), ("conv0_1.conv1", 192, 64, 512, 3, 1), ("conv1_0.conv2", 128, 128, 256, 3, 1),
    ("conv1_1.conv1", 384, 128, 256, 3, 1), ("conv2_1.conv1", 512, 256, 128, 3, 1), ("conv3_1.conv1", 768, 384, 64, 3, 1),
    ("conv4_1.conv1", 1024, 512, 32, 3, 1), ("conv5_0.conv2", 768, 768, 16, 3, 1), ("short1_1 1x1", 384, 128, 256, 1, 1),
    ("D.block1 s2", 64, 64, 512, 3, 2), ("D.block3 s2", 128, 128, 256, 3, 2), ("spade gb L0", 8, 128, 512, 3, 1), ("x2map L0", 64, 8, 512, 3, 1),
    ("D.conv0", 8, 64, 512, 3, 1), ("spade gb L1", 8, 256, 256, 3, 1), ("conv0_0 dgradsplit", 192, 64, 512, 3, 1),
    ("final 1x1", 64, 8, 512, 1, 1), ("short0_1 1x1", 192, 64, 512, 1, 1), ("short2_1 1x1", 768, 256, 128, 1, 1),
]
which = sys.argv[1:] or ["fwd", "dgrad", "wgrad"]
ONLY = os.environ.get("ONLY")
for name, cin, cout, hw, k, stride in SHAPES:
    if ONLY and ONLY not in name:
        continue
    x = ops.empty_nhwc(B, cin, hw, hw, torch.bfloat16); x.normal_()
    w = torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k)
    oh = (hw + 2 * (k // 2) - k) // stride + 1
    y = ops.empty_nhwc(B, cout, oh, oh, torch.bfloat16)
    dy = ops.empty_nhwc(B, cout, oh, oh, torch.bfloat16); dy.normal_()
    dx = ops.empty_nhwc(B, cin, hw, hw, torch.bfloat16)
    dw = torch.empty_like(w)
    fl = 2.0 * B * oh * oh * cin * cout * k * k
    fns = {"fwd": lambda: conv_tc.forward(x, w, None, y, stride, k // 2, 0, 0.0),
           "dgrad": lambda: conv_tc.dgrad(dy, w, dx, stride, k // 2),
           "wgrad": lambda: conv_tc.wgrad(x, dy, dw, stride, k // 2)}
    out = []
    for nm in which:
        f = fns[nm]
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out.append("%s %7.3f ms %6.0f TF/s" % (nm, ms, fl / ms / 1e9))
    print("%-16s cin %4d cout %4d hw %3d k%d s%d | %s" % (name, cin, cout, hw, k, stride, " | ".join(out)), flush=True)
    del x, y, dy, dx
