import sys, math
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F
import ssunet_gan_b200 as ssg
from ssunet_gan_b200 import ops, models_seg_gan
import ssunet_oracle as O
ssg.set_compute_dtype(torch.float32); ssg.set_conv_impl("simt")
xd,_=O.synthetic_batch(3,3,96,96,seed=5)
d=models_seg_gan.Discriminator(3); d.load_state_dict(O.portable_state_dict(O.discriminator_spec(3))); d.cuda().train()
t=ops.to_nhwc(xd.cuda())
for blk in d.conv_blocks: t=blk(t)
flat=ops.adaptive_avg_pool_flat(t,6,6)
lo=d.fc2(d.fc1(flat,act=ops.ACT_LEAKY,slope=0.2))
ops.bce_with_logits_const(lo,1.0).backward()
torch.cuda.synchronize()
print("done", float(d.fc2.weight.grad.abs().sum()))
