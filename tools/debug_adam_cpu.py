import sys, os, numpy as np, torch
sys.path.insert(0,''+os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'')
from oracle import sg2_oracle as orc
import torch.nn.functional as F
g=dict(np.load(''+os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/golden/sg2_layer8.npz'))
from rewriting_b200.utils.stylegan2 import SeqStyleGAN2
model = orc.seeded_state_dict(lambda: SeqStyleGAN2(256, style_dim=512, n_mlp=8, mconv='seq'))
sd=model.state_dict()
W0=sd['layer8.sconv.mconv.dconv.weight'].clone()
k=torch.from_numpy(g['goal_in_fmap']); st=torch.from_numpy(g['goal_in_style']); tgt=torch.from_numpy(g['goal_out_fmap'])
nw=sd['layer8.sconv.noise.weight']; bias=sd['layer8.sconv.activate.bias']; d=torch.from_numpy(g['d'])
N=30
W_ref=orc.insert_loop(W0,k,st,tgt,nw,bias,d,N,piter=10,lr=0.05)
def manual(omb1, omb2):
    w=W0.clone().requires_grad_(True)
    with torch.no_grad(): ortho=w-orc.projected_conv(w,d)
    m=torch.zeros_like(w); v=torch.zeros_like(w)
    f=np.float32
    for it in range(N):
        loss=F.l1_loss(tgt, orc.target_forward(k,st,w,nw,bias,True))
        w.grad=None; loss.backward(); gr=w.grad
        with torch.no_grad():
            m = m + (gr-m)*float(f(omb1))
            v = v*float(f(0.999)) + float(f(omb2))*gr*gr
            step=it+1
            bc1=1-0.9**step; bc2=1-0.999**step
            ss=float(f(0.05/bc1)); b2s=float(f(bc2**0.5))
            denom=v.sqrt()/b2s+1e-8
            w -= ss*(m/denom)
            if it%10==0 or it==N-1:
                w[...]=ortho+orc.projected_conv(w,d)
    return w.detach()
f=np.float32
for name,(a,b) in {'double-rounded':(1-0.9,1-0.999),'float-arith':(float(f(1)-f(0.9)),float(f(1)-f(0.999)))}.items():
    Wm=manual(a,b)
    diff=(Wm-W_ref).abs()
    print(name,'max diff',diff.max().item(),'entries>1e-4',int((diff>1e-4).sum()),'>1e-6',int((diff>1e-6).sum()))
