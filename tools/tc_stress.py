"""Repeat the 26 000 x 1204 tcgen05 forward + weight-gradient GEMMs and count runs outside the 1e-5 bar (the
race hunt of profiles/README.md s10).  python tools/tc_stress.py [iterations]; GSAGE_LIB selects a build."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "graphsage-simple_b200"))
import torch
from graphsage import ops
n,k_in,d_out,act=26000,1204,128,1
g=torch.Generator(device="cuda").manual_seed(n+k_in)
x=ops.empty_rows(n,k_in,"cuda"); x.copy_(torch.randn(n,k_in,device="cuda",generator=g))
w=torch.randn(d_out,k_in,device="cuda",generator=g)/k_in**0.5
gh=torch.randn(n,d_out,device="cuda",generator=g)
h=torch.empty((n,d_out),device="cuda"); ops.encoder_fwd_tc(x,w,act,h)
ref_h=torch.relu(x.double()@w.double().t())
hd=h.double(); dz=gh.double()*(hd>0).double(); ref=(dz.t()@x.double())
bad_f=bad_w=0; worst=0
iters=int(sys.argv[1]) if len(sys.argv)>1 else 60
for it in range(iters):
    h2=torch.full((n,d_out),float('nan'),device='cuda'); ops.encoder_fwd_tc(x,w,act,h2)
    ef=float((h2.double()-ref_h).abs().max()/ref_h.abs().max())
    gw=torch.full((d_out,k_in),float('nan'),device='cuda'); ops.encoder_wgrad_tc(x,h,gh,act,gw)
    e=float((gw.double()-ref).abs().max()/ref.abs().max())
    bad_f+= (not ef<1e-5); bad_w += (not e<1e-5); worst=max(worst,e)
print(os.environ.get("GSAGE_LIB","default"), "fwd bad", bad_f, "wgrad bad", bad_w, "of", iters, "worst", worst)
