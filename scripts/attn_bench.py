"""Attention kernel timings (CUDA events, L2 flushed between iterations) for the path's three big shapes at bs=64:
the mma.sync kernel (impl 1) against the tcgen05/TMEM kernel (impl 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops

SHAPES = [("dino", 128, 261, 16, 16, 64, False), ("siglip", 128, 256, 16, 16, 72, False),
          ("qwen", 64, 625, 14, 2, 64, True)]


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    global SHAPES
    if only:
        SHAPES = [s for s in SHAPES if s[0] == only]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, B, S, H, HKV, hd, causal in SHAPES:
        qkv = (torch.randn(B * S, (H + 2 * HKV) * hd, device="cuda") * 1.0).to(torch.bfloat16)
        flops = 4.0 * B * H * S * S * hd * (0.5 if causal else 1.0)
        res = {}
        for impl in (1, 2):
            ops.set_attention_impl(impl)
            for _ in range(3):
                out = ops.attention(qkv, B, S, H, HKV, hd, causal)
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = ops.attention(qkv, B, S, H, HKV, hd, causal)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            res[impl] = (ts[len(ts) // 2], out)
        ops.set_attention_impl(0)
        d = (res[1][1].float() - res[2][1].float()).abs().max().item()
        print(f"{name:7s} B={B} S={S} hd={hd} causal={causal}: mma.sync {res[1][0]*1e3:8.1f} us ({flops/res[1][0]/1e9:6.1f} TF/s)"
              f" | tcgen05 {res[2][0]*1e3:8.1f} us ({flops/res[2][0]/1e9:6.1f} TF/s) | max|diff| {d:.4f}")


if __name__ == "__main__":
    main()
