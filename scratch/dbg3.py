import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn as nn, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
torch.manual_seed(0)
def chain(cfgs, hw):
    mine=[models_seg_gan.ConvolutionalBlock(ci,co,3,stride=s,batch_norm=bn,activation='LeakyReLu') for (ci,co,s,bn) in cfgs]
    ref=[]
    for m,(ci,co,s,bn) in zip(mine,cfgs):
        layers=[nn.Conv2d(ci,co,3,s,1)]
        if bn: layers.append(nn.BatchNorm2d(co))
        layers.append(nn.LeakyReLU(0.2))
        r=nn.Sequential(*layers); r[0].load_state_dict(m.conv_block[0].state_dict())
        ref.append(r)
    x=torch.randn(3,cfgs[0][0],hw,hw)
    xr=x.clone().requires_grad_(True); t=xr
    for r in ref: t=r(t)
    gy=torch.randn(t.shape); t.backward(gy)
    xc=x.cuda().requires_grad_(True); u=xc
    mine=[m.cuda() for m in mine]
    for m in mine: u=m(u)
    u.backward(gy.cuda())
    print(cfgs, "y %.2e dx %.2e"%(rel(u,t),rel(xc.grad,xr.grad)))
    for i,(m,r) in enumerate(zip(mine,ref)):
        print("   blk",i,"dW %.2e db %.2e"%(rel(m.conv_block[0].weight.grad,r[0].weight.grad), rel(m.conv_block[0].bias.grad,r[0].bias.grad)), ("dgamma %.2e dbeta %.2e"%(rel(m.conv_block[1].weight.grad,r[1].weight.grad),rel(m.conv_block[1].bias.grad,r[1].bias.grad))) if len(r)==3 else "")
chain([(64,64,1,True),(64,64,2,True)],12)
chain([(64,64,1,True),(64,64,1,True)],12)
chain([(64,64,1,False),(64,64,2,True)],12)
chain([(64,64,1,True),(64,64,2,False)],12)
chain([(8,8,1,True),(8,8,1,True)],6)
