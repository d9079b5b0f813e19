"""Find the first module of the fp32 (SIMT) generator whose output differs between repeated identical forwards."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import ssunet_oracle as O
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import models_seg_gan

ssg.set_compute_dtype(torch.float32)
ssg.set_conv_impl("simt")
g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False})
g.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net.")))
g = g.cuda().train()
x, _ = O.synthetic_batch(2, 3, 64, 64, seed=1234)
x = x.cuda()
names = {m: n for n, m in g.named_modules()}
rec = []
def hook(m, inp, out):
    o = out[0] if isinstance(out, (tuple, list)) else out
    if torch.is_tensor(o):
        rec.append((names[m], o.detach().float().clone()))
for m in g.modules():
    if len(list(m.children())) == 0 or type(m).__name__ in ("BasicBlock", "SPADE"):
        m.register_forward_hook(hook)
ref = None
junk = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    rec.clear()
    junk.append(torch.randn(1 << (18 + it % 5), device="cuda"))      # perturb the allocator between runs
    if it % 3 == 2:
        junk.clear()
    with torch.no_grad():
        out = g(x)
    torch.cuda.synchronize()
    cur = list(rec)
    if ref is None:
        ref = cur
        print("modules hooked:", len(ref))
        continue
    bad = [(i, n) for i, ((n, a), (_, b)) in enumerate(zip(ref, cur)) if not torch.equal(a, b)]
    if bad:
        i, n = bad[0]
        a, b = ref[i][1], cur[i][1]
        d = (a - b).abs()
        print("run %d: %d outputs differ; first = #%d %s  max|d| %.3e  n_diff %d / %d  rel %.2e ; logits rel %.2e" % (
            it, len(bad), i, n, float(d.max()), int((d > 0).sum()), d.numel(), float((a - b).norm() / b.norm()),
            float((ref[-1][1] - cur[-1][1]).norm() / ref[-1][1].norm())))
        idx = torch.nonzero(d > 0)[:5]
        print("   first differing indices:", idx.tolist(), "prev module:", ref[i - 1][0] if i else None)
    else:
        print("run %d: identical" % it)
