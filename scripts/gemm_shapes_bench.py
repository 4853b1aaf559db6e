"""Sustained throughput of the tcgen05 GEMM on the engine's bs=64 shapes WITH their real epilogues
(python scripts/gemm_shapes_bench.py [seconds per shape]); cuBLAS (torch.matmul, no epilogue) beside it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vla_adapter_b200 import ops

SEC = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
# name, M, N, K, bias, act, colscale, resid
SHAPES = [
    ("dino.qkv", 33408, 3072, 1024, 1, "none", 0, 0),
    ("dino.proj", 33408, 1024, 1024, 1, "none", 1, 1),
    ("dino.fc1", 33408, 4096, 1024, 1, "gelu", 0, 0),
    ("dino.fc2", 33408, 1024, 4096, 1, "none", 1, 1),
    ("sig.qkv", 32768, 3456, 1152, 1, "none", 0, 0),
    ("sig.proj", 32768, 1152, 1152, 1, "none", 0, 1),
    ("sig.fc1", 32768, 4304, 1152, 1, "gelu", 0, 0),
    ("sig.fc2", 32768, 1152, 4304, 1, "none", 0, 1),
    ("proj.fc1", 32768, 8704, 2176, 1, "gelu", 0, 0),
    ("proj.fc2", 32768, 896, 8704, 1, "gelu", 0, 0),
    ("llm.qkv", 40000, 1152, 896, 1, "none", 0, 0),
    ("llm.o", 40000, 896, 896, 0, "none", 0, 1),
    ("llm.gateup", 40000, 9728, 896, 0, "swiglu", 0, 0),
    ("llm.down", 40000, 896, 4864, 0, "none", 0, 1),
    ("pol.kv_vis", 32768, 1792, 896, 1, "none", 1, 0),
]


def timeit(fn, sec):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    while True:
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if ms >= sec * 1e3:
            return ms / n
        n = max(n * 2, int(n * sec * 1e3 / max(ms, 1e-3) * 1.1))


print(f"{'shape':12s} {'M':>6s} {'N':>5s} {'K':>5s} epi            ours(bn) TF/s                      cuBLAS TF/s")
for name, M, N, K, hb, act, hs, hr in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    n_out = N // 2 if act == "swiglu" else N
    out = torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, device="cuda") if hb else None
    cs = (torch.rand(N, device="cuda") + 0.5) if hs else None
    res = {}
    for bn in (0, 256, 192, 128):
        if act == "swiglu" and bn == 64:
            continue
        fn = lambda: ops.linear(a, w, bias=bias, act=act, colscale=cs, resid=out if hr else None, out=out, force_bn=bn)
        res[bn] = 2.0 * M * N * K / (timeit(fn, SEC) * 1e-3) / 1e12
        if hr:
            out.zero_()
    wt = w.T
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cb = 2.0 * M * N * K / (timeit(lambda: torch.matmul(a, wt, out=o2), SEC) * 1e-3) / 1e12
    epi = ("b" if hb else "-") + act[:4] + ("s" if hs else "-") + ("r" if hr else "-")
    print(f"{name:12s} {M:6d} {N:5d} {K:5d} {epi:10s} auto {res[0]:6.0f} | 256 {res[256]:6.0f} | 192 {res[192]:6.0f} | 128 {res[128]:6.0f} | cublas {cb:6.0f}", flush=True)
