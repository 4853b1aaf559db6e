"""Tile-width sweep of the bs=1 GEMM shapes with COLD weights (a ring of weight copies larger than L2, as in a real
forward where 2.7 GB of weights pass between two uses of a matrix): python scripts/bs1_tile_sweep.py
Prints the average time per launch (CUDA events around the ring) for the heuristic's choice and every forced width."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops

SHAPES = [("dino.qkv", 522, 3072, 1024, "none", False), ("dino.proj", 522, 1024, 1024, "none", True),
          ("dino.fc1", 522, 4096, 1024, "gelu", False), ("dino.fc2", 522, 1024, 4096, "none", True),
          ("sig.qkv", 512, 3456, 1152, "none", False), ("sig.fc1", 512, 4304, 1152, "gelu", False),
          ("sig.fc2", 512, 1152, 4304, "none", True), ("llm.qkv", 625, 1152, 896, "none", False),
          ("llm.o", 625, 896, 896, "none", True), ("llm.gateup", 625, 9728, 896, "swiglu", False),
          ("llm.down", 625, 896, 4864, "none", True)]
dev = torch.device("cuda")
for name, M, N, K, act, resid in SHAPES:
    n_copies = max(4, int(160e6 // (N * K * 2)) + 1)
    ws = [(torch.randn(N, K, device=dev) * K ** -0.5).bfloat16() for _ in range(n_copies)]
    a = torch.randn(M, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev) if act != "swiglu" else None
    n_out = N // 2 if act == "swiglu" else N
    out = torch.zeros(M, n_out, device=dev, dtype=torch.bfloat16)
    res = []
    for bn in (0, 64, 128, 192, 224, 256):
        if act == "swiglu" and bn not in (0, 128, 256):
            continue
        try:
            for w in ws[:2]:
                ops.linear(a, w, bias=bias, act=act, resid=out if resid else None, out=out, force_bn=bn)
            torch.cuda.synchronize()
            # the Python wrapper + tensor-map encode cost ~20 us per launch on the host: capture the ring into a CUDA
            # graph so that the device time is what is measured
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for w in ws:
                    ops.linear(a, w, bias=bias, act=act, resid=out if resid else None, out=out, force_bn=bn)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for rep in range(3):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            res.append(f"bn={bn or 'auto'}: {e0.elapsed_time(e1) / (3 * n_copies) * 1e3:6.2f} us")
        except Exception as ex:
            res.append(f"bn={bn}: {str(ex).splitlines()[0][:40]}")
    print(f"{name:11s} M={M} N={N} K={K}  " + "  ".join(res), flush=True)
