import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn as nn, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan, _lib
import ssunet_gan_b200.ops as opsmod
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
mode=sys.argv[1]
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
xd,_=O.synthetic_batch(3,3,96,96,seed=5)
sd=O.portable_state_dict(O.discriminator_spec(3)); O._leafify(sd)
lo=O.discriminator(sd,xd,True); ks=O.trainable_keys(sd)
ref=dict(zip(ks, torch.autograd.grad(F.binary_cross_entropy_with_logits(lo,torch.ones_like(lo)),[sd[k] for k in ks])))
log=[]
orig=_lib.call
def traced(name,*args,**kw):
    log.append((name, torch.cuda.current_stream().cuda_stream, [a.data_ptr() if isinstance(a,torch.Tensor) else None for a in args]))
    if 'sync' in mode: torch.cuda.synchronize()
    orig(name,*args,**kw)
    if 'sync' in mode: torch.cuda.synchronize()
opsmod.call=traced
d=models_seg_gan.Discriminator(3); d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.cuda().train()
xc=xd.cuda()
t=ops.to_nhwc(xc)
for blk in d.conv_blocks: t=blk(t)
flat=ops.adaptive_avg_pool_flat(t,6,6)
lo=d.fc2(d.fc1(flat,act=ops.ACT_LEAKY,slope=0.2))
ops.bce_with_logits_const(lo,1.0).backward()
torch.cuda.synchronize()
g={k:p.grad for k,p in d.named_parameters()}
print(mode, "blk6.W %.2e"%rel(g["conv_blocks.6.conv_block.0.weight"],ref["conv_blocks.6.conv_block.0.weight"]), "streams", set(l[1] for l in log), "ncalls", len(log))
# aliasing check: any call whose output ptr equals one of its own inputs
for name,s,ptrs in log:
    ps=[p for p in ptrs if p]
    if len(ps)!=len(set(ps)): print("ALIAS in", name, ptrs)
