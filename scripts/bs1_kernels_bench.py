"""Back-to-back device time of the bs=1 kernels (CUDA events around 200 launches each, warm): what one launch costs
when the forward is a chain of ~550 dependent small kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, M, N, K, act in [("dino.qkv", 522, 3072, 1024, "none"), ("dino.proj", 522, 1024, 1024, "none"),
                           ("dino.fc2", 522, 1024, 4096, "none"), ("llm.qkv", 625, 1152, 896, "none"),
                           ("llm.o", 625, 896, 896, "none"), ("llm.gateup", 625, 9728, 896, "swiglu"),
                           ("llm.down", 625, 896, 4864, "none"), ("pol.q", 8, 896, 896, "none"),
                           ("pol.kv_vis", 512, 1792, 896, "none")]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    out = torch.empty(M, N // 2 if act == "swiglu" else N, device="cuda", dtype=torch.bfloat16)
    res = []
    for bn in (0, 64, 128, 256):
        if act == "swiglu" and bn == 64:
            res.append(float("nan")); continue
        res.append(timeit(lambda: ops.linear(a, w, act=act, out=out, force_bn=bn)))
    print(f"gemm {name:11s} M={M:4d} N={N:5d} K={K:5d}: auto {res[0]:6.1f} us | bn64 {res[1]:6.1f} | bn128 {res[2]:6.1f} | bn256 {res[3]:6.1f}")
x = torch.randn(522, 1024, device="cuda").bfloat16(); wv = torch.randn(1024, device="cuda")
print(f"layernorm 522x1024: {timeit(lambda: ops.layernorm(x, wv, wv, 1e-6)):6.1f} us")
for name, B, S, H, HKV, hd, causal in [("dino", 2, 261, 16, 16, 64, False), ("siglip", 2, 256, 16, 16, 72, False), ("qwen", 1, 625, 14, 2, 64, True)]:
    qkv = torch.randn(B * S, (H + 2 * HKV) * hd, device="cuda").bfloat16()
    r = []
    for impl in (1, 2):
        ops.set_attention_impl(impl)
        r.append(timeit(lambda: ops.attention(qkv, B, S, H, HKV, hd, causal)))
    ops.set_attention_impl(0)
    print(f"attention {name:7s}: mma.sync {r[0]:6.1f} us | tcgen05 {r[1]:6.1f} us")
