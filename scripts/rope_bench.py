import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops
M,N,K,S=40000,1152,896,625
a=torch.randn(M,K,device="cuda").bfloat16(); w=(torch.randn(N,K,device="cuda")*K**-0.5).bfloat16(); bias=torch.randn(N,device="cuda")
cos=torch.randn(S,32,device="cuda"); sin=torch.randn(S,32,device="cuda")
def t(fn,n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
out=torch.empty(M,N,device="cuda",dtype=torch.bfloat16)
for bn in (0,128,192,256):
    print("plain gemm bn",bn, t(lambda: ops.linear(a,w,bias=bias,out=out,force_bn=bn)))
print("rope kernel", t(lambda: ops.rope_(out,0,16,64,S,1e6)))
print("fused", t(lambda: ops.linear_rope(a,w,bias,cos,sin,1024,S)))
