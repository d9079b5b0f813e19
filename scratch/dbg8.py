import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan, losses, _lib
import ssunet_oracle as O
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
xd,_=O.synthetic_batch(3,3,96,96,seed=5)
d=models_seg_gan.Discriminator(3); d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.cuda().train()
xc=xd.cuda()
# poison the allocator's free memory with NaNs
junk=torch.full((768*1024*1024,), float('nan'), device='cuda'); del junk
# trace calls: after each C-ABI call check all float tensor args for NaN (sync) -- report first offender
orig_call=_lib.call
import ssunet_gan_b200.ops as opsmod
def traced(name,*args,**kw):
    orig_call(name,*args,**kw)
    torch.cuda.synchronize()
    for i,a in enumerate(args):
        if isinstance(a,torch.Tensor) and a.is_floating_point():
            if bool(torch.isnan(a).any()):
                print("NaN after", name, "arg", i, tuple(a.shape), a.dtype, "count", int(torch.isnan(a).sum()))
opsmod.call=traced
lo=d(xc)
ops.bce_with_logits_const(lo,1.0).backward()
torch.cuda.synchronize()
for k,p in d.named_parameters():
    if torch.isnan(p.grad).any(): print("NaN grad", k)
print("D done")
