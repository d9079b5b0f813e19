"""Discriminator head fc1 (18432 -> 1024, batch 16) forward / backward timings (CUDA events, L2 flushed between calls by a 256 MB write)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops

ssg.set_compute_dtype(torch.bfloat16)
m, k, n = 16, 18432, 1024
x = torch.randn(m, k, device="cuda").bfloat16().requires_grad_(True)
w = (torch.randn(n, k, device="cuda") / 100).requires_grad_(True)
b = torch.randn(n, device="cuda").requires_grad_(True)
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")


def timed(fn, iters=10):
    ms = 0.0
    for _ in range(3):
        fn()
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / iters


y = ops.linear(x, w, b, ops.ACT_LEAKY, 0.2)
ref = torch.nn.functional.leaky_relu(x.float() @ w.t() + b, 0.2)
# raw kernel timing: 30 back-to-back launches rotating over 3 weight copies (226 MB > L2), one event pair around all of them
ws = [w.detach().clone() for _ in range(3)]
yb = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
xd = x.detach()
def raw():
    for i in range(30):
        ops.call("ssg_linear_fwd", xd, ws[i % 3], b.detach(), yb, ops.dtype_code(torch.bfloat16), m, k, n, ops.ACT_LEAKY, 0.2, None)
raw(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); raw(); e1.record(); torch.cuda.synchronize()
print("raw fwd kernel %.4f ms per launch" % (e0.elapsed_time(e1) / 30))
dx32 = torch.empty(m, k, device="cuda")
def rawd():
    for i in range(30):
        ops.call("ssg_linear_dgrad", yb, ws[i % 3], dx32, ops.dtype_code(torch.bfloat16), m, k, n, None)
rawd(); torch.cuda.synchronize()
e0.record(); rawd(); e1.record(); torch.cuda.synchronize()
print("raw dgrad kernel %.4f ms per launch" % (e0.elapsed_time(e1) / 30))
dw = torch.empty(n, k, device="cuda"); db = torch.empty(n, device="cuda")
def raww():
    for i in range(30):
        ops.call("ssg_linear_wgrad", xd, yb, dw, db, ops.dtype_code(torch.bfloat16), m, k, n)
raww(); torch.cuda.synchronize()
e0.record(); raww(); e1.record(); torch.cuda.synchronize()
print("raw wgrad kernel %.4f ms per launch" % (e0.elapsed_time(e1) / 30))
print("fwd rel err", float((y.float() - ref).norm() / ref.norm()))
print("fwd  %.4f ms" % timed(lambda: ops.linear(x, w, b, ops.ACT_LEAKY, 0.2)))
gy = torch.randn_like(y)


def fb():
    x.grad = w.grad = b.grad = None
    ops.linear(x, w, b, ops.ACT_LEAKY, 0.2).backward(gy)


print("fwd+bwd %.4f ms" % timed(fb))
