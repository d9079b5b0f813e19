import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F, numpy as np
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan, losses
import ssunet_oracle as O
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
gm=models_seg_gan.Generator({"arch":"UNet_R_SS_v2","num_classes":3,"input_channels":3,"deep_supervision":False})
gm.load_state_dict(O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net."))); gm.cuda().train()
x,t=O.synthetic_batch(2,3,64,64,seed=1234)
out=gm(x.cuda()); loss=losses.BCEDiceLoss()(out,t.cuda()); loss.backward()
sd=O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net.")); O._leafify(sd)
oo=O.unet_r_ss_v2(sd,x,True,prefix="net."); ol=O.bce_dice_loss(oo,t); keys=O.trainable_keys(sd)
og=dict(zip(keys, torch.autograd.grad(ol,[sd[k] for k in keys])))
names=[k for k,_ in gm.named_parameters()]
order=["final","SPADE0_1","conv0_1","SPADE1_1","conv1_1","SPADE2_1","conv2_1","conv_head3_1","SPADE3_1","conv3_1","conv_head4_1","SPADE4_1","conv4_1","conv_head5_0","SPADE5_0","conv5_0","SPADE4_0","conv4_0","SPADE3_0","conv3_0","SPADE2_0","conv2_0","SPADE1_0","conv1_0","SPADE0_0","conv0_0"]
g={k:p.grad for k,p in gm.named_parameters()}
for o in order:
    ks=[k for k in names if k.startswith("net."+o+".")]
    print(o, " ".join("%s=%.1e"%(k.split(o+".")[1], rel(g[k],og[k])) for k in ks if float(og[k].abs().sum())>1e-3))
