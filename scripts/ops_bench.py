"""Achieved HBM bandwidth of the bandwidth-bound kernels at their bs=64 shapes (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for name, M, D, rms in [("layernorm dino", 33408, 1024, False), ("layernorm siglip", 32768, 1152, False), ("rmsnorm llm", 40000, 896, True)]:
    x = torch.randn(M, D, device="cuda").bfloat16()
    w = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
    ms = timeit((lambda: ops.rmsnorm(x, w, 1e-6)) if rms else (lambda: ops.layernorm(x, w, b, 1e-6)))
    print(f"{name:18s} M={M} D={D}: {ms*1e3:7.1f} us  {2*M*D*2/ms/1e6:7.0f} GB/s (read+write)")
x = torch.randn(40000, 1152, device="cuda").bfloat16()
ms = timeit(lambda: ops.rope_(x, 0, 16, 64, 625, 1e6))
print(f"{'rope llm':18s} M=40000 heads=16: {ms*1e3:7.1f} us  {2*40000*1024*2/ms/1e6:7.0f} GB/s (read+write of q,k)")
