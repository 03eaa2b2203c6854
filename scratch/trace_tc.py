import sys, os; sys.path.insert(0,'/root/repo/graphsage-simple_b200')
import torch
tr = torch.zeros(64*16, dtype=torch.int64, device='cuda')
os.environ['GSAGE_TC_TRACE'] = str(tr.data_ptr())
from graphsage import ops
n,k_in,d=25275,1204,128
g = torch.Generator(device="cuda").manual_seed(1)
x = ops.empty_rows(n, k_in, "cuda"); x.copy_(torch.randn(n, k_in, device="cuda", generator=g))
w = torch.randn(d, k_in, device="cuda", generator=g) / k_in ** 0.5
h = torch.empty((n, d), device="cuda")
for it in range(3): ops.encoder_fwd_tc(x, w, 1, h)
torch.cuda.synchronize()
t = tr.cpu().view(64,16)
t0 = int(t[0,0])
names=['P:empty','P:issued','Y:full','Y:done','X:full','X:afree','X:stored','M:ready','M:issued']
print('chunk '+' '.join('%9s'%n for n in names))
for c in range(38):
    print('%5d '%c + ' '.join('%9d'%(int(t[c,e])-t0) for e in range(9)))
