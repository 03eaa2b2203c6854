import sys; sys.path.insert(0,'/root/repo/graphsage-simple_b200')
import torch, numpy as np
from graphsage import ops
def relerr(a,b): 
    a=a.double(); b=b.double(); return float((a-b).abs().max()/b.abs().max())
def tf32(t):
    b = t.view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)
for n,k_in in [(4000,1204),(4000,256),(4000,64)]:
    d=128
    g = torch.Generator(device="cuda").manual_seed(n + k_in)
    x = tf32(torch.randn(n, k_in, device="cuda", generator=g)).contiguous()
    w = tf32(torch.randn(d, k_in, device="cuda", generator=g) / k_in ** 0.5).contiguous()
    h = torch.empty((n, d), device="cuda")
    ops.encoder_fwd_tc(x, w, 0, h)
    ref = x.double() @ w.double().t()
    err = (h.double()-ref)
    print(n,k_in,'exact-input relerr', relerr(h,ref), 'mean signed err*sign(ref)', float((err*ref.sign()).mean()), 'mean abs err', float(err.abs().mean()), 'mean|ref|', float(ref.abs().mean()))
    # positive-only data to expose RZ bias
    x = tf32(torch.rand(n, k_in, device="cuda", generator=g)).contiguous()
    w = tf32(torch.rand(d, k_in, device="cuda", generator=g)).contiguous()
    ops.encoder_fwd_tc(x, w, 0, h)
    ref = x.double() @ w.double().t()
    err = (h.double()-ref)
    print('   positive data: relerr', relerr(h,ref), 'mean rel signed', float((err/ref).mean()), ' fp32 torch:', relerr((x@w.t()), ref))
