import sys; sys.path.insert(0,'/root/repo/graphsage-simple_b200')
import torch
from graphsage import ops
torch.manual_seed(0)
n,k_in,d=256,128,128
x=torch.ones(n,k_in,device='cuda'); 
h=torch.ones(n,d,device='cuda'); gh=torch.ones(n,d,device='cuda')
ws=torch.full((ops.encoder_wgrad_tc_ws_floats(n,k_in,d),),-5.0,device='cuda')
gw=torch.full((d,k_in),float('nan'),device='cuda')
ops.encoder_wgrad_tc(x,h,gh,0,gw,ws=ws)
torch.cuda.synchronize()
print('gw',gw[:2,:8], gw.unique()[:10])
print('dz_hi', ws[:8], 'dz_lo', ws[n*d:n*d+8])
part=ws[2*n*d:]
print('part', part[:8], part.unique()[:10], part.numel())
# structured: x[r,c]=c, dz=1 -> gw[m,c]=n*c
x=torch.arange(k_in,device='cuda').float().repeat(n,1).contiguous()
ops.encoder_wgrad_tc(x,h,gh,0,gw,ws=ws); torch.cuda.synchronize()
print('gw row0', gw[0,:40])
gh=torch.arange(d,device='cuda').float().repeat(n,1).contiguous()
x=torch.ones(n,k_in,device='cuda')
ops.encoder_wgrad_tc(x,h,gh,0,gw,ws=ws); torch.cuda.synchronize()
print('gw col0', gw[:40,0])
