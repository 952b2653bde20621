import os, sys, torch
sys.path.insert(0, "/root/repo")
from tsasr_b200 import ops
B,T,U,H,V=16,400,100,640,1000
dev=torch.device("cuda:0"); g=torch.Generator().manual_seed(0)
enc=(0.5*torch.randn(B,T,H,generator=g)).bfloat16().to(dev); dec=(0.5*torch.randn(B,U,H,generator=g)).bfloat16().to(dev)
W=((torch.rand(V,H,generator=g)*2-1)/H**0.5).bfloat16().to(dev); b=((torch.rand(V,generator=g)*2-1)/H**0.5).to(dev)
tg=torch.randint(1,V,(B,U-1),generator=g,dtype=torch.int32).to(dev)
ll=torch.full((B,),T,dtype=torch.int32).to(dev); tl=torch.full((B,),U-1,dtype=torch.int32).to(dev)
for _ in range(3): ops.joint_fwd(enc,dec,W,b,tg,ll,tl,0,0,0.01)
torch.cuda.synchronize()
