"""Per-call device-time breakdown of one eager G+D step (CUDA events around every C-ABI call), grouped by entry point and
shape.  python scratch/prof_step.py [batch] [size]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import _lib, models_seg_gan, optim, train_step
import ssunet_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ssg.set_compute_dtype(torch.bfloat16)
torch.manual_seed(41)
g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False}).cuda().train()
d = models_seg_gan.Discriminator(3).cuda().train()
og = optim.FusedClampAdam(g.parameters(), lr=2e-5)
od = optim.FusedClampAdam(d.parameters(), lr=2e-5)
x, t = O.synthetic_batch(B, 3, S, S, seed=1234)
x, t = x.cuda(), t.cuda()
for _ in range(2):
    train_step.gan_train_step(g, d, og, od, x, t, with_metrics=False)
torch.cuda.synchronize()

events = []
orig = _lib.call


def call(name, *args, flops=0.0):
    key = tuple(a for a in args if isinstance(a, (int,)) and not isinstance(a, bool))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(name, *args, flops=flops)
    e1.record()
    events.append((name, key, flops, e0, e1))


_lib.call = call
import ssunet_gan_b200.ops as ops_mod, ssunet_gan_b200.conv_tc as ctc
for m in list(sys.modules.values()):
    if m is not None and getattr(m, "__name__", "").startswith("ssunet_gan_b200") and getattr(m, "call", None) is orig:
        m.call = call
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
train_step.gan_train_step(g, d, og, od, x, t, with_metrics=False)
s1.record()
torch.cuda.synchronize()
tot = s0.elapsed_time(s1)
agg = collections.OrderedDict()
for name, key, fl, e0, e1 in events:
    k = (name, key)
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl
inside = sum(a[1] for a in agg.values())
print("eager step %.2f ms; inside C-ABI calls %.2f ms (%d calls)" % (tot, inside, len(events)))
byname = collections.defaultdict(float)
for (name, key), a in agg.items():
    byname[name] += a[1]
for name, ms in sorted(byname.items(), key=lambda kv: -kv[1])[:25]:
    print("  %-32s %8.3f ms %5.1f%%" % (name, ms, 100 * ms / inside))
print()
TOP = int(os.environ.get("TOP", "70"))
for (name, key), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:TOP]:
    tf = " %6.0f TF/s" % (a[2] / a[1] / 1e9) if a[2] else ""
    print("%-26s n=%2d %8.3f ms%s  %s" % (name, a[0], a[1], tf, key))
