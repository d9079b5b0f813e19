import sys; sys.path.insert(0,'oracle')
import torch, torch.nn.functional as F
import ssunet_oracle as O
torch.set_num_threads(8)
def q(t, on=True): return t.bfloat16().float() if on else t
def run(size, cfg, batch=2):
    sd=O.portable_state_dict(O.unet_r_ss_v2_spec(3,3,prefix="net."))
    x,_=O.synthetic_batch(batch,3,size,size,seed=1234)
    P="net."
    def conv(t,w,b=None,pad=1,qw=True): return F.conv2d(t,q(w,qw and cfg['w']),b,1,pad)
    def bn(p,t,relu,res=None):
        y=F.batch_norm(t,None,None,sd[p+".weight"],sd[p+".bias"],True,0.1,1e-5)
        if res is not None: y=y+res
        if relu: y=F.relu(y)
        return q(y,cfg['act'])
    def block(p,t):
        c1=q(conv(t,sd[p+".conv1.weight"]),cfg['conv_out'])
        r1=bn(p+".bn1",c1,True)
        c2=q(conv(r1,sd[p+".conv2.weight"]),cfg['conv_out'])
        sc=q(conv(t,sd[p+".shortcut.0.weight"],None,0),cfg['conv_out'])
        return bn(p+".bn2",c2,True,sc)
    def spade(p,t):
        seg=q(conv(t,sd[p+".x2map.weight"],sd[p+".x2map.bias"]),cfg['seg'])
        a=q(F.relu(conv(seg,sd[p+".mlp_shared.0.weight"],sd[p+".mlp_shared.0.bias"],qw=cfg['seg'])),cfg['actv'])
        g=q(conv(a,sd[p+".mlp_gamma.weight"],sd[p+".mlp_gamma.bias"]),cfg['gb'])
        b=q(conv(a,sd[p+".mlp_beta.weight"],sd[p+".mlp_beta.bias"]),cfg['gb'])
        return q(t*(1+g)+b,cfg['act'])
    def stage(c,s,t): return spade(P+s,block(P+c,t))
    x0=q(x,cfg['act'])
    e0=stage("conv0_0","SPADE0_0",x0); p0,_=F.max_pool2d(e0,2,2,return_indices=True)
    e1=stage("conv1_0","SPADE1_0",p0); p1,_=F.max_pool2d(e1,2,2,return_indices=True)
    e2=stage("conv2_0","SPADE2_0",p1); p2,i2=F.max_pool2d(e2,2,2,return_indices=True)
    e3=stage("conv3_0","SPADE3_0",p2); p3,i3=F.max_pool2d(e3,2,2,return_indices=True)
    e4=stage("conv4_0","SPADE4_0",p3); p4,i4=F.max_pool2d(e4,2,2,return_indices=True)
    e5=stage("conv5_0","SPADE5_0",p4); e5=q(conv(e5,sd[P+"conv_head5_0.weight"],None,0),cfg['act'])
    d4=stage("conv4_1","SPADE4_1",torch.cat([e4,F.max_unpool2d(e5,i4,2,2)],1)); d4=q(conv(d4,sd[P+"conv_head4_1.weight"],None,0),cfg['act'])
    d3=stage("conv3_1","SPADE3_1",torch.cat([e3,F.max_unpool2d(d4,i3,2,2)],1)); d3=q(conv(d3,sd[P+"conv_head3_1.weight"],None,0),cfg['act'])
    d2=stage("conv2_1","SPADE2_1",torch.cat([e2,F.max_unpool2d(d3,i2,2,2)],1))
    up=lambda t: q(F.interpolate(t,scale_factor=2,mode="bilinear",align_corners=True),cfg['act'])
    d1=stage("conv1_1","SPADE1_1",torch.cat([e1,up(d2)],1))
    d0=stage("conv0_1","SPADE0_1",torch.cat([e0,up(d1)],1))
    return F.conv2d(d0,q(sd[P+"final.weight"],cfg['w']),sd[P+"final.bias"])
def rel(a,b): return float((a-b).norm()/b.norm())
none=dict(w=False,act=False,conv_out=False,seg=False,actv=False,gb=False)
with torch.no_grad():
    for size in (64,128):
        ref=run(size,none)
        for name,cfg in [("all bf16",dict(w=True,act=True,conv_out=True,seg=True,actv=True,gb=True)),
                         ("seg+actv fp32",dict(w=True,act=True,conv_out=True,seg=False,actv=False,gb=True)),
                         ("seg fp32 only",dict(w=True,act=True,conv_out=True,seg=False,actv=True,gb=True)),
                         ("seg,actv,gb fp32",dict(w=True,act=True,conv_out=True,seg=False,actv=False,gb=False)),
                         ("conv_out fp32 (BN sees fp32)",dict(w=True,act=True,conv_out=False,seg=True,actv=True,gb=True)),
                         ("conv_out+seg+actv fp32",dict(w=True,act=True,conv_out=False,seg=False,actv=False,gb=True)),
                         ("conv_out+seg+actv+gb fp32",dict(w=True,act=True,conv_out=False,seg=False,actv=False,gb=False)),
                         ("only weights bf16",dict(w=True,act=False,conv_out=False,seg=False,actv=False,gb=False)),
                         ("only act bf16",dict(w=False,act=True,conv_out=False,seg=False,actv=False,gb=False))]:
            print(size, "%-34s rel %.3e"%(name, rel(run(size,cfg),ref)))
