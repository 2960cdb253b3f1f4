import sys; sys.path.insert(0,'/root/repo')
import torch
from prompt_tts_b200 import ops
dev='cuda'
B,L,D,V=32,550,768,150
g=torch.Generator(device=dev).manual_seed(0)
ids=torch.randint(1,V,(B,L),device=dev,generator=g).to(torch.int32)
lens=torch.randint(100,L+1,(B,),device=dev,generator=g)
ids=ids*(torch.arange(L,device=dev)[None,:]<lens[:,None]).to(torch.int32)
dy=torch.randn(B,L,D,device=dev,generator=g).to(torch.bfloat16)
dE=torch.zeros(V,D,device=dev)
def f(): ops.call("text_embed_bwd", ops._p(ids), ops._p(dy), ops._p(dE), B, L, D, V, ops._stream())
f(); torch.cuda.synchronize()
ref=torch.zeros(V,D,device=dev).index_add_(0, ids.reshape(-1).long(), dy.reshape(-1,D).float())
print("rel err", ((dE-ref).norm()/ref.norm()).item())
gr=torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(10): f()
gr.replay(); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print("us per call", e0.elapsed_time(e1)/10*1e3)
