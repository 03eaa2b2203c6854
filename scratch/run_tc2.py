import sys; sys.path.insert(0,'/root/repo/graphsage-simple_b200')
import torch
from graphsage import ops
d=128
for n,k_in in [(25275,1204),(25275,608),(25275,128),(25275,32),(18944,1204),(12800,1204)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = ops.empty_rows(n, k_in, "cuda"); x.copy_(torch.randn(n, k_in, device="cuda", generator=g))
    w = torch.randn(d, k_in, device="cuda", generator=g) / k_in ** 0.5
    h = torch.empty((n, d), device="cuda")
    for it in range(3): ops.encoder_fwd_tc(x, w, 1, h)
    torch.cuda.synchronize()
    e0,e1=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e0.record()
    for it in range(20): ops.encoder_fwd_tc(x, w, 1, h)
    e1.record(); torch.cuda.synchronize()
    print(n,k_in,'tiles',(n+127)//128,'chunks',(k_in+31)//32,'fwd_tc us', 1000*e0.elapsed_time(e1)/20)
