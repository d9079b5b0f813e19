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
routs=[]; r=xd
for i in range(8):
    p="conv_blocks.%d.conv_block"%i
    r=F.conv2d(r,sd[p+".0.weight"],sd[p+".0.bias"],1 if i%2==0 else 2,1)
    if i: r=O.batch_norm(sd,p+".1",r,True)
    r=F.leaky_relu(r,0.2); r.retain_grad(); routs.append(r)
rf=F.adaptive_avg_pool2d(r,(6,6)).reshape(3,-1)
rl=F.linear(F.leaky_relu(F.linear(rf,sd["fc1.weight"],sd["fc1.bias"]),0.2),sd["fc2.weight"],sd["fc2.bias"])
F.binary_cross_entropy_with_logits(rl,torch.ones_like(rl)).backward()
shapes=[tuple(o.shape) for o in routs]
bufs=[torch.zeros(s,device='cuda') for s in shapes]
fbufs=[torch.zeros(s,device='cuda') for s in shapes]
d=models_seg_gan.Discriminator(3); d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.cuda().train()
xc=xd.cuda()
t=ops.to_nhwc(xc)
def mk(i):
    def h(g): bufs[i].copy_(g)
    return h
for i,blk in enumerate(d.conv_blocks):
    t=blk(t); fbufs[i].copy_(t); t.register_hook(mk(i))
flat=ops.adaptive_avg_pool_flat(t,6,6)
lo=d.fc2(d.fc1(flat,act=ops.ACT_LEAKY,slope=0.2))
ops.bce_with_logits_const(lo,1.0).backward()
torch.cuda.synchronize()
for i in range(7,-1,-1):
    p="conv_blocks.%d.conv_block"%i
    m=d.conv_blocks[i].conv_block
    s="block %d out %.2e gout %.2e dW %.2e"%(i, rel(fbufs[i],routs[i]), rel(bufs[i],routs[i].grad), rel(m[0].weight.grad, sd[p+".0.weight"].grad))
    if i: s+=" dgamma %.2e dbeta %.2e"%(rel(m[1].weight.grad, sd[p+".1.weight"].grad), rel(m[1].bias.grad, sd[p+".1.bias"].grad))
    print(s)
