import sys; sys.path.insert(0,'/root/repo/graphsage-simple_b200')
import torch, numpy as np
from graphsage import ops
def relerr(a,b): 
    a=a.double(); b=b.double(); return float((a-b).abs().max()/b.abs().max())
for n,k_in in [(26000,1204),(26000,256),(4000,1204),(12800,1204), (26000,1204)]:
    d=128
    g = torch.Generator(device="cuda").manual_seed(n + k_in)
    x = ops.empty_rows(n, k_in, "cuda"); x.copy_(torch.randn(n, k_in, device="cuda", generator=g))
    w = torch.randn(d, k_in, device="cuda", generator=g) / k_in ** 0.5
    gh = torch.randn(n, d, device="cuda", generator=g)
    h = torch.full((n, d), float("nan"), device="cuda")
    ops.encoder_fwd_tc(x, w, 1, h)
    ref = torch.relu(x.double() @ w.double().t())
    e1 = relerr(h, ref)
    bad = ((h.double()-ref).abs() > 1e-4*ref.abs().max()).nonzero()
    gw = torch.full((d, k_in), float("nan"), device="cuda")
    ops.encoder_wgrad_tc(x, h, gh, 1, gw)
    dz = gh.double() * (h.double() > 0)
    refw = dz.t() @ x.double()
    e2 = relerr(gw, refw)
    badw = ((gw.double()-refw).abs() > 1e-4*refw.abs().max()).nonzero()
    print(n,k_in,'fwd',e1, 'nbad', bad.shape[0], bad[:5].tolist(), 'wgrad', e2, 'nbad', badw.shape[0], badw[:5].tolist(), 'nan', torch.isnan(h).sum().item(), torch.isnan(gw).sum().item())
