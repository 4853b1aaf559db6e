"""Runs one engine GEMM shape a few times (for `ncu --set full -k regex:gemm_bf16 -s 3 -c 1`): python scripts/gemm_one.py NAME"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops
SHAPES = {"dino.fc1": (33408, 4096, 1024, "gelu"), "dino.proj": (33408, 1024, 1024, "none"), "llm.gateup": (40000, 9728, 896, "swiglu"),
          "llm.o": (40000, 896, 896, "none"), "proj.fc1": (32768, 8704, 2176, "gelu")}
M, N, K, act = SHAPES[sys.argv[1]]
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda") if act != "swiglu" else None
for _ in range(6):
    out = ops.linear(a, w, bias=bias, act=act)
torch.cuda.synchronize()
