import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan, losses
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
g=torch.Generator().manual_seed(1)
# BN bwd C=512 rows=108 leaky
for shape in [(3,512,6,6),(3,512,12,12),(2,64,64,64)]:
    x=torch.randn(shape,generator=g)*0.5; gamma=1+0.1*torch.randn(shape[1],generator=g); beta=0.1*torch.randn(shape[1],generator=g)
    xr=x.clone().requires_grad_(True); gr=gamma.clone().requires_grad_(True); br=beta.clone().requires_grad_(True)
    yr=F.leaky_relu(F.batch_norm(xr,None,None,gr,br,True,0.1,1e-5),0.2); gy=torch.randn(shape,generator=g); yr.backward(gy)
    xc=x.cuda().requires_grad_(True); gc=gamma.cuda().requires_grad_(True); bc=beta.cuda().requires_grad_(True)
    y=ops.batch_norm(xc,gc,bc,torch.zeros(shape[1]).cuda(),torch.ones(shape[1]).cuda(),True,0.1,1e-5,None,ops.ACT_LEAKY,0.2); y.backward(gy.cuda())
    print("BN", shape, "y %.2e dx %.2e dg %.2e db %.2e"%(rel(y,yr),rel(xc.grad,xr.grad),rel(gc.grad,gr.grad),rel(bc.grad,br.grad)))
# conv s2 512->512 12x12
for (n,cin,cout,h,w,k,s,p) in [(3,512,512,12,12,3,2,1),(3,256,512,12,12,3,1,1),(2,128,64,64,64,3,1,1)]:
    x=torch.randn(n,cin,h,w,generator=g); wt=torch.randn(cout,cin,k,k,generator=g)/math.sqrt(cin*k*k); b=torch.randn(cout,generator=g)
    xr=x.clone().requires_grad_(True); wr=wt.clone().requires_grad_(True); br=b.clone().requires_grad_(True)
    yr=F.conv2d(xr,wr,br,s,p); gy=torch.randn(yr.shape,generator=g); yr.backward(gy)
    xc=x.cuda().requires_grad_(True); wc=wt.cuda().requires_grad_(True); bc=b.cuda().requires_grad_(True)
    y=ops.conv2d(xc,wc,bc,s,p); y.backward(gy.cuda())
    print("conv",(n,cin,cout,h,w,k,s,p),"y %.2e dx %.2e dw %.2e db %.2e"%(rel(y,yr),rel(xc.grad,xr.grad),rel(wc.grad,wr.grad),rel(bc.grad,br.grad)))
# D in backward order with double-precision oracle
sdd=O.portable_state_dict(O.discriminator_spec(3))
sdd64={k:(v.double() if v.is_floating_point() else v) for k,v in sdd.items()}
O._leafify(sdd); O._leafify(sdd64)
xd,_=O.synthetic_batch(3,3,96,96,seed=5)
def run(sd,x):
    lo=O.discriminator(sd,x,True); l=F.binary_cross_entropy_with_logits(lo,torch.ones_like(lo)); ks=O.trainable_keys(sd)
    return dict(zip(ks, torch.autograd.grad(l,[sd[k] for k in ks])))
g32=run(sdd,xd); g64=run(sdd64,xd.double())
d=models_seg_gan.Discriminator(3); d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.cuda().train()
lo=d(xd.cuda()); l=ops.bce_with_logits_const(lo,1.0); l.backward()
for k,p in d.named_parameters():
    if 'conv_block.0.bias' in k and 'blocks.0' not in k: continue
    print("  ours-vs-f64 %.3e  cpu32-vs-f64 %.3e  %s" % (rel(p.grad,g64[k]), rel(g32[k],g64[k]), k))
