import sys; sys.path.insert(0,'/root/repo/graphsage-simple_b200')
import torch
from graphsage import ops
n,k_in,d=25275,1204,128
g = torch.Generator(device="cuda").manual_seed(1)
x = ops.empty_rows(n, k_in, "cuda"); x.copy_(torch.randn(n, k_in, device="cuda", generator=g))
w = torch.randn(d, k_in, device="cuda", generator=g) / k_in ** 0.5
gh = torch.randn(n, d, device="cuda", generator=g)
h = torch.empty((n, d), device="cuda")
gw = torch.empty((d, k_in), device="cuda")
ws1 = torch.empty(ops.encoder_fwd_tc_ws_floats(k_in,d), device='cuda')
ws2 = torch.empty(ops.encoder_wgrad_tc_ws_floats(n,k_in,d), device='cuda')
for it in range(3):
    ops.encoder_fwd_tc(x, w, 1, h, ws=ws1)
    ops.encoder_wgrad_tc(x, h, gh, 1, gw, ws=ws2)
torch.cuda.synchronize()
e0,e1,e2=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record()
for it in range(10): ops.encoder_fwd_tc(x, w, 1, h, ws=ws1)
e1.record()
for it in range(10): ops.encoder_wgrad_tc(x, h, gh, 1, gw, ws=ws2)
e2.record(); torch.cuda.synchronize()
print('fwd_tc ms', e0.elapsed_time(e1)/10, 'wgrad_tc ms', e1.elapsed_time(e2)/10)
