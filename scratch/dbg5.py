import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn as nn, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
xd,_=O.synthetic_batch(3,3,96,96,seed=5)
sd=O.portable_state_dict(O.discriminator_spec(3)); O._leafify(sd)
lo=O.discriminator(sd,xd,True); ks=O.trainable_keys(sd)
ref=dict(zip(ks, torch.autograd.grad(F.binary_cross_entropy_with_logits(lo,torch.ones_like(lo)),[sd[k] for k in ks])))
def ours(req_in, retain, via_forward):
    d=models_seg_gan.Discriminator(3); d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.cuda().train()
    xc=xd.cuda()
    if req_in: xc.requires_grad_(True)
    if via_forward:
        lo=d(xc)
    else:
        t=ops.to_nhwc(xc)
        for blk in d.conv_blocks:
            t=blk(t)
            if retain: t.retain_grad()
        flat=ops.adaptive_avg_pool_flat(t,6,6)
        lo=d.fc2(d.fc1(flat,act=ops.ACT_LEAKY,slope=0.2))
    ops.bce_with_logits_const(lo,1.0).backward()
    torch.cuda.synchronize()
    g={k:p.grad for k,p in d.named_parameters()}
    print("req_in",req_in,"retain",retain,"fwd",via_forward, " blk6.W %.2e blk3.W %.2e blk0.W %.2e"%(rel(g["conv_blocks.6.conv_block.0.weight"],ref["conv_blocks.6.conv_block.0.weight"]), rel(g["conv_blocks.3.conv_block.0.weight"],ref["conv_blocks.3.conv_block.0.weight"]), rel(g["conv_blocks.0.conv_block.0.weight"],ref["conv_blocks.0.conv_block.0.weight"])))
ours(True,False,True)
ours(False,False,True)
ours(False,False,False)
ours(False,True,False)
ours(True,True,False)
ours(True,False,True)
