"""One traced launch of the tcgen05 attention kernel (VLA_FA_TRACE) for a shape; prints CTA 0's event log."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vla_adapter_b200 import ops
name, B, S, H, HKV, hd, causal = {"dino": ("dino", 128, 261, 16, 16, 64, False), "siglip": ("siglip", 128, 256, 16, 16, 72, False),
                                  "qwen": ("qwen", 64, 625, 14, 2, 64, True)}[sys.argv[1]]
qkv = torch.randn(B * S, (H + 2 * HKV) * hd, device="cuda").to(torch.bfloat16)
ops.set_attention_impl(2)
for _ in range(2):
    ops.attention(qkv, B, S, H, HKV, hd, causal)
torch.cuda.synchronize()
os.environ["VLA_FA_TRACE"] = sys.argv[2]
ops.attention(qkv, B, S, H, HKV, hd, causal)
torch.cuda.synchronize()
