import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan, losses, archs
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
x,t=O.synthetic_batch(2,3,64,64,seed=1234)
# ---- oracle with intermediates of conv0_1
keep={}
orig_bb=O.basic_block
def bb(sd,p,xx,training=True,sync_stats=None):
    if p!="net.conv0_1": return orig_bb(sd,p,xx,training,sync_stats)
    c1=F.conv2d(xx,sd[p+".conv1.weight"],None,1,1); c1.retain_grad()
    r1=F.relu(O.batch_norm(sd,p+".bn1",c1,training)); r1.retain_grad()
    c2=F.conv2d(r1,sd[p+".conv2.weight"],None,1,1); c2.retain_grad()
    b2=O.batch_norm(sd,p+".bn2",c2,training)
    sc=F.conv2d(xx,sd[p+".shortcut.0.weight"])
    keep.update(c1=c1,r1=r1,c2=c2)
    return F.relu(b2+sc)
O.basic_block=bb
sd=O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net.")); O._leafify(sd)
oo=O.unet_r_ss_v2(sd,x,True,prefix="net."); O.bce_dice_loss(oo,t).backward()
# ---- ours with hooks into preallocated buffers
gm=models_seg_gan.Generator({"arch":"UNet_R_SS_v2","num_classes":3,"input_channels":3,"deep_supervision":False})
gm.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net."))); gm.cuda().train()
buf={k:torch.zeros(2,64,64,64,device='cuda') for k in ("c1","r1","c2","gc1","gr1","gc2")}
blk=gm.net.conv0_1
def fwd(xx):
    xx=ops.to_nhwc(xx)
    c1=blk.conv1(xx); buf["c1"].copy_(c1); c1.register_hook(lambda g: (buf["gc1"].copy_(g), None)[1])
    r1=blk.bn1(c1,act=ops.ACT_RELU); buf["r1"].copy_(r1); r1.register_hook(lambda g: (buf["gr1"].copy_(g), None)[1])
    c2=blk.conv2(r1); buf["c2"].copy_(c2); c2.register_hook(lambda g: (buf["gc2"].copy_(g), None)[1])
    sc=blk.shortcut[0](xx)
    return blk.bn2(c2,residual=sc,act=ops.ACT_RELU)
blk.forward=fwd
out=gm(x.cuda()); losses.BCEDiceLoss()(out,t.cuda()).backward(); torch.cuda.synchronize()
print("fwd: c1 %.2e r1 %.2e c2 %.2e"%(rel(buf["c1"],keep["c1"]),rel(buf["r1"],keep["r1"]),rel(buf["c2"],keep["c2"])))
print("bwd: gc2 %.2e gr1 %.2e gc1 %.2e"%(rel(buf["gc2"],keep["c2"].grad),rel(buf["gr1"],keep["r1"].grad),rel(buf["gc1"],keep["c1"].grad)))
print("params: "+" ".join("%s=%.1e"%(k,rel(p.grad,sd["net.conv0_1."+k].grad)) for k,p in blk.named_parameters()))
dz=(keep["r1"].grad*(keep["r1"]>0)); dzo=(buf["gr1"].cpu()*(buf["r1"].cpu()>0))
print("sum dz rel %.2e ; mask mismatches %d of %d ; gr1 max abs %.3e; diff at mismatches"%(rel(dzo.sum((0,2,3)),dz.sum((0,2,3))), int(((buf["r1"].cpu()>0)!=(keep["r1"]>0)).sum()), dz.numel(), float(keep["r1"].grad.abs().max())))
print("zero fraction in ref r1: %.3f ; count exact-zero bn output sign issues"%float((keep["r1"]==0).float().mean()))
