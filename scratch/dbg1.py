import sys, math, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
for dt in (torch.float32, torch.bfloat16):
    ssg.set_compute_dtype(dt); ssg.set_conv_impl("simt")
    for (n,cin,cout,h,w,k,stride,pad,act) in [(2,64,64,16,16,3,1,1,0),(2,64,64,16,16,3,1,1,2),(2,3,64,20,24,3,1,1,2),(2,5,7,9,11,3,2,1,2)]:
        g=torch.Generator().manual_seed(1)
        x=torch.randn(n,cin,h,w,generator=g); wt=torch.randn(cout,cin,k,k,generator=g)/math.sqrt(cin*k*k)
        xr=x.clone().requires_grad_(True); wr=wt.clone().requires_grad_(True)
        yr=F.conv2d(xr,wr,None,stride,pad)
        if act: yr=F.leaky_relu(yr,0.2)
        gy=torch.randn(yr.shape,generator=g); yr.backward(gy)
        xc=x.cuda().requires_grad_(True); wc=wt.cuda().requires_grad_(True)
        y=ops.conv2d(xc,wc,None,stride,pad,act,0.2); y.backward(gy.cuda().to(dt))
        # reference with rounded operands
        xq=x.to(dt).float().requires_grad_(True); wq=wt.to(dt).float().requires_grad_(True)
        yq=F.conv2d(xq,wq,None,stride,pad)
        if act: yq=F.leaky_relu(yq,0.2)
        yq.backward(gy.to(dt).float())
        print(dt, (cin,cout,k,stride,act), "fwd %.2e dx %.2e dw %.2e | vs rounded-operand ref: fwd %.2e dx %.2e dw %.2e" % (rel(y.float(),yr), rel(xc.grad,xr.grad), rel(wc.grad,wr.grad), rel(y.float(),yq), rel(xc.grad,xq.grad), rel(wc.grad,wq.grad)))
# generator grads per key in fp32
z=np.load('tests/golden/generator_fwd_bwd_2x64.npz')
from ssunet_gan_b200 import models_seg_gan, losses
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
gm=models_seg_gan.Generator({"arch":"UNet_R_SS_v2","num_classes":3,"input_channels":3,"deep_supervision":False})
gm.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net."))); gm.cuda().train()
x,t=O.synthetic_batch(2,3,64,64,seed=1234)
out=gm(x.cuda()); loss=losses.BCEDiceLoss()(out,t.cuda()); loss.backward()
# oracle grads
sd=O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net.")); O._leafify(sd)
oo=O.unet_r_ss_v2(sd,x,True,prefix="net."); ol=O.bce_dice_loss(oo,t); keys=O.trainable_keys(sd)
og=dict(zip(keys, torch.autograd.grad(ol,[sd[k] for k in keys])))
rows=[]
for k,p in gm.named_parameters():
    rows.append((rel(p.grad, og[k]), k, float(og[k].abs().sum())))
rows.sort(reverse=True)
print("fp32 G: logits rel %.2e" % rel(out, oo))
for r in rows[:25]: print("  %.3e %s abs-sum %.3e" % r)
# discriminator
zd=np.load('tests/golden/discriminator_fwd_bwd_3x96.npz')
d=models_seg_gan.Discriminator(3); d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.cuda().train()
xd,_=O.synthetic_batch(3,3,96,96,seed=5); xc=xd.cuda().requires_grad_(True)
lo=d(xc); l=ops.bce_with_logits_const(lo,1.0); l.backward()
sdd=O.portable_state_dict(O.discriminator_spec(3)); O._leafify(sdd)
xo=xd.clone().requires_grad_(True); loo=O.discriminator(sdd,xo,True); lo_=F.binary_cross_entropy_with_logits(loo,torch.ones_like(loo))
dk=O.trainable_keys(sdd); dg=dict(zip(dk, torch.autograd.grad(lo_,[sdd[k] for k in dk])))
print("fp32 D logit rel %.2e" % rel(lo, loo))
for k,p in d.named_parameters():
    print("  %.3e %s abs-sum ref %.3e ours %.3e" % (rel(p.grad, dg[k]), k, float(dg[k].abs().sum()), float(p.grad.abs().sum())))
# bf16 generator at 64 and 128 vs oracle
for size in (64,128,256):
    x,t=O.synthetic_batch(1 if size==256 else 2,3,size,size,seed=1234)
    sd=O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net."))
    with torch.no_grad(): oo=O.unet_r_ss_v2(sd,x,True,prefix="net.")
    for dt in (torch.float32, torch.bfloat16):
        ssg.set_compute_dtype(dt)
        gm.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net.")))
        with torch.no_grad(): out=gm(x.cuda())
        print("G fwd size", size, dt, "rel %.3e" % rel(out,oo))
