"""Eager (no CUDA graph) pipelined steps with CUDA events on both streams -> per-op timeline."""
import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/graphsage-simple_b200')
import numpy as np, torch, torch.nn as nn
import bench
from graphsage import ops, sampling
from graphsage.aggregators import MeanAggregator
from graphsage.encoders import Encoder
from graphsage.graph import CSRGraph
from graphsage.model import SupervisedGraphSage
from graphsage.engine import engine_for
class A: nodes=233000; pairs=5800000; feat=602; hidden=128; classes=41; k1=10; k2=25
args=A()
dev=torch.device('cuda')
rowptr,col=bench.build_graph_arrays(args.nodes,args.pairs)
graph=CSRGraph(rowptr,col,dev)
table=ops.empty_rows(args.nodes,args.feat,dev,zero=True); table.copy_(torch.randn(args.nodes,args.feat,device=dev))
emb=nn.Embedding(args.nodes,args.feat,device='meta'); emb.weight=nn.Parameter(table,requires_grad=False)
def build():
    agg1=MeanAggregator(emb,cuda=True); enc1=Encoder(emb,args.feat,128,graph,agg1,num_sample=10,gcn=False,cuda=True)
    agg2=MeanAggregator(lambda n: enc1(n).t(),cuda=True); enc2=Encoder(lambda n: enc1(n).t(),128,128,graph,agg2,num_sample=25,base_model=enc1,gcn=False,cuda=True)
    return SupervisedGraphSage(41,enc2)
model=build()
B=1024
eng=engine_for(model,B); eng.use_graphs=False; eng.enable_pipeline()
rng=np.random.default_rng(0)
nodes=torch.from_numpy(rng.integers(0,args.nodes,(16,B)).astype(np.int32)).to(dev)
labels=torch.from_numpy(rng.integers(0,41,(16,B))).to(dev)
eng.stage_device(nodes[0],labels[0],1); eng.prime(B)
events=[]
orig={}
def wrap(name):
    fn=getattr(ops,name); orig[name]=fn
    def timed(*a,**kw):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); out=fn(*a,**kw); e1.record()
        events.append((name+('[L1]' if kw.get('n_dev') is not None else ''), torch.cuda.current_stream()==eng._side, e0,e1)); return out
    setattr(ops,name,timed)
for nm in ("sample_csr","dedup_remap","gather_mean_fwd","encoder_fwd","encoder_fwd_tc","classifier_xent","encoder_bwd","encoder_wgrad_tc","scatter_mean_bwd","sgd_step"): wrap(nm)
for i in range(6):
    eng.stage_device(nodes[i+1],labels[i+1],i+2,slot=1-eng.cur)
    events.clear()
    t0=torch.cuda.Event(enable_timing=True); t0.record()
    eng.train_step_pipelined(B,0.01,B)
    t1=torch.cuda.Event(enable_timing=True); t1.record()
    torch.cuda.synchronize()
print('step total ms', t0.elapsed_time(t1))
for name,side,e0,e1 in events:
    print('%-24s %-5s start %7.1f  end %7.1f  dur %6.1f us'%(name,'side' if side else 'main', 1000*t0.elapsed_time(e0), 1000*t0.elapsed_time(e1), 1000*e0.elapsed_time(e1)))
