"""Quick stand-alone GEMM throughput probe (not a pytest): python tests/gemm_bench_gpu.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops

shapes = [(33408, 3072, 1024), (33408, 4096, 1024), (33408, 1024, 4096), (32768, 4304, 1152), (32768, 1152, 4304),
          (40000, 1152, 896), (40000, 9728, 896), (40000, 896, 4864), (32768, 8704, 2176), (8192, 8192, 8192)]
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = {}
    for bn in (256, 128):
        for _ in range(3):
            ops.linear(a, w, out=out, force_bn=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.linear(a, w, out=out, force_bn=bn)
        e1.record(); torch.cuda.synchronize()
        res[bn] = 2 * M * N * K * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    for _ in range(3):
        torch.matmul(a, w.T, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(a, w.T, out=out)
    e1.record(); torch.cuda.synchronize()
    cb = 2 * M * N * K * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print(f"M={M} N={N} K={K}: bn256 {res[256]:.0f} TF/s  bn128 {res[128]:.0f} TF/s  cublas {cb:.0f} TF/s", flush=True)
