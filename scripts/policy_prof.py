#!/usr/bin/env python
"""Phase timeline of the fused small-batch policy kernel (VLA_POLICY_PROF=1): prints clock64 deltas per block."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VLA_POLICY_PROF"] = "1"
import torch
import bench
from vla_adapter_b200 import tokens
from vla_adapter_b200.engine import VLAEngine
from vla_adapter_b200.weights import load_random_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pro = len(sys.argv) > 2 and sys.argv[2] == "pro"
eng = VLAEngine(n_images=2, pro=pro, max_batch=8, max_prompt_len=48, device=0)
load_random_weights(eng, seed=0, n_images=2, action_dim=7, proprio_dim=8, pro=pro)
eng.finalize()
dev = torch.device("cuda", 0)
pix, ids, prop = bench.synth_inputs(B, 48, seed=0, device=dev)
ext, _, _, aq, _ = tokens.build(ids.cpu(), None, 7)
ext_d, aq_d = ext.to(dev), aq.to(dev)
for _ in range(20):
    eng.predict_device(pix, ext_d, aq_d, prop)
torch.cuda.synchronize()
eng.close()
