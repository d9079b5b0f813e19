import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, archs
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
torch.manual_seed(0)
blk=archs.BasicBlock(192,64)
sd={k:v.clone() for k,v in blk.state_dict().items()}
x=torch.randn(2,192,64,64).abs()   # positive-mean input like post-ReLU/concat activations
gy=torch.randn(2,64,64,64)
# reference in fp64
sd64={k:(v.double().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k,v in sd.items()}
xr=x.double().requires_grad_(True)
c1=F.conv2d(xr,sd64["conv1.weight"],None,1,1); c1.retain_grad()
b1=F.batch_norm(c1,None,None,sd64["bn1.weight"],sd64["bn1.bias"],True,0.1,1e-5)
r1=F.relu(b1); r1.retain_grad()
c2=F.conv2d(r1,sd64["conv2.weight"],None,1,1); c2.retain_grad()
b2=F.batch_norm(c2,None,None,sd64["bn2.weight"],sd64["bn2.bias"],True,0.1,1e-5)
sc=F.conv2d(xr,sd64["shortcut.0.weight"])
out=F.relu(b2+sc); out.backward(gy.double())
blk.cuda().train()
xc=x.cuda().requires_grad_(True)
o=blk(xc); o.backward(gy.cuda())
print("out %.2e dx %.2e"%(rel(o,out),rel(xc.grad,xr.grad)))
for k,p in blk.named_parameters(): print("  %-20s %.2e"%(k, rel(p.grad, sd64[k].grad)))
# isolated pieces fed with reference tensors
dy_r1=r1.grad.float(); c1f=c1.detach().float(); r1f=r1.detach().float()
mean=c1f.mean((0,2,3)); var=c1f.var((0,2,3),unbiased=False); inv=1/torch.sqrt(var+1e-5)
xq=ops.to_nhwc(c1f.cuda()); yq=ops.to_nhwc(r1f.cuda()); dq=ops.to_nhwc(dy_r1.cuda())
sums=torch.empty(128,dtype=torch.float64,device='cuda')
ops.call("ssg_bn_bwd_reduce", dq,yq,xq,0,2*64*64,64,mean.cuda(),inv.cuda(),1,0.0,sums)
dz=(r1.grad*(r1>0)).detach()
xhat=((c1.detach()-c1.detach().mean((0,2,3),keepdim=True))/torch.sqrt(c1.detach().var((0,2,3),unbiased=False,keepdim=True)+1e-5))
print("isolated reduce: sum dz %.2e  sum dz*xhat %.2e"%(rel(sums[:64],dz.sum((0,2,3))), rel(sums[64:],(dz*xhat).sum((0,2,3)))))
print("cancellation: |sum dz| / sum|dz| = %.2e"%float(dz.sum((0,2,3)).abs().sum()/dz.abs().sum()))
dxq=ops.empty_nhwc(2,64,64,64,torch.float32)
ops.call("ssg_bn_bwd_apply", dq,yq,xq,dxq,None,0,2*64*64,64,mean.cuda(),inv.cuda(),sd["bn1.weight"].cuda(),sums,float(2*64*64),1,0.0,1)
print("isolated bn dx %.2e"%rel(dxq, c1.grad))
# conv2 dgrad isolated with reference dy
wp=torch.empty(64*64*9,device='cuda'); ops.call("ssg_pack_conv_weight", sd["conv2.weight"].cuda().contiguous(), wp, 0, 1, 64,64,3,3, None)
d2=ops.to_nhwc(c2.grad.float().cuda()); dxo=ops.empty_nhwc(2,64,64,64,torch.float32)
ops.call("ssg_conv2d_dgrad_simt", d2, wp, dxo, 0, 2,64,64,64,64,3,3,1,1)
print("isolated conv2 dgrad %.2e"%rel(dxo, r1.grad))
