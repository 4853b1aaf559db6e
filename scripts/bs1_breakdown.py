#!/usr/bin/env python
"""Where a bs=1 forward spends its time: per-subsystem device time of an eager call, the graph-replayed latency, and
the per-launch GEMM list (VLA_GEMM_PROF_CSV) aggregated by shape."""
import collections, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSV = "/tmp/bs1_gemm.csv"
if os.path.exists(CSV):
    os.remove(CSV)
os.environ["VLA_GEMM_PROF_CSV"] = CSV
import torch
import bench
from vla_adapter_b200 import _lib, tokens
from vla_adapter_b200.engine import VLAEngine
from vla_adapter_b200.weights import load_random_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = VLAEngine(n_images=2, pro=False, max_batch=8, max_prompt_len=48, device=0)
load_random_weights(eng, seed=0, n_images=2, action_dim=7, proprio_dim=8, pro=False)
eng.finalize()
lib = _lib.load()
dev = torch.device("cuda", 0)
pix, ids, prop = bench.synth_inputs(B, 48, seed=0, device=dev)
ext, _, _, aq, _ = tokens.build(ids.cpu(), None, 7)
ext_d, aq_d = ext.to(dev), aq.to(dev)
step = lambda: eng.predict_device(pix, ext_d, aq_d, prop)
for _ in range(10):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    step()
e1.record()
torch.cuda.synchronize()
print(f"graph replay, device-resident inputs: {e0.elapsed_time(e1) / 50:.3f} ms per call (B={B})")
seg3 = (C.c_float * 3)()
lib.vla_segment_timing(eng._h, 1)
step(); step()
torch.cuda.synchronize()
_lib.check(lib.vla_segment_times(eng._h, seg3), eng._h)
lib.vla_segment_timing(eng._h, 0)
print(f"eager segments: towers+projector {seg3[0]:.3f} ms, llm {seg3[1]:.3f} ms, policy {seg3[2]:.3f} ms")
lib.vla_profile_gemm(1)
step()
torch.cuda.synchronize()
g_ms, g_n = C.c_double(0), C.c_longlong(0)
lib.vla_profile_gemm_read(C.byref(g_ms), C.byref(g_n))
lib.vla_profile_gemm(0)
print(f"GEMM launches {g_n.value}, summed event time {g_ms.value:.3f} ms")
agg = collections.OrderedDict()
for line in open(CSV):
    f = line.strip().split(",")
    key = tuple(f[:6])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += float(f[6]); a[2] += float(f[8])
print("rows,batches,N,K,bn,act: launches, avg event us, avg CTA-0 lifetime us")
for k, (n, ms, us) in agg.items():
    print(f"  {','.join(k):32s} {n:4d}  {ms / n * 1e3:7.2f}  {us / n:7.2f}")
eng.close()
