import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn as nn, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
sd=O.portable_state_dict(O.discriminator_spec(3))
d=models_seg_gan.Discriminator(3); d.load_state_dict(sd); d.cuda().train()
xd,_=O.synthetic_batch(3,3,96,96,seed=5)
# ours with retained grads on block outputs
outs=[]; t=ops.to_nhwc(xd.cuda())
for blk in d.conv_blocks:
    t=blk(t); t.retain_grad(); outs.append(t)
flat=ops.adaptive_avg_pool_flat(t,6,6); flat.retain_grad()
hid=d.fc1(flat,act=ops.ACT_LEAKY,slope=0.2); lo=d.fc2(hid)
l=ops.bce_with_logits_const(lo,1.0); l.backward()
# reference
O._leafify(sd)
routs=[]; r=xd
for i in range(8):
    p="conv_blocks.%d.conv_block"%i
    r=F.conv2d(r,sd[p+".0.weight"],sd[p+".0.bias"],1 if i%2==0 else 2,1)
    if i: r=O.batch_norm(sd,p+".1",r,True)
    r=F.leaky_relu(r,0.2); r.retain_grad(); routs.append(r)
rf=F.adaptive_avg_pool2d(r,(6,6)).reshape(3,-1); rf.retain_grad()
rl=F.linear(F.leaky_relu(F.linear(rf,sd["fc1.weight"],sd["fc1.bias"]),0.2),sd["fc2.weight"],sd["fc2.bias"])
F.binary_cross_entropy_with_logits(rl,torch.ones_like(rl)).backward()
print("flat grad %.2e"%rel(flat.grad, rf.grad))
for i in range(7,-1,-1):
    print("block",i,"out %.2e grad-of-out %.2e"%(rel(outs[i],routs[i]), rel(outs[i].grad, routs[i].grad)), tuple(outs[i].shape))
