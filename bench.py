#!/usr/bin/env python
"""bench.py - action chunks/sec of the batched predict_action path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our engine, N ranks x B samples (weak scaling)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A step is one pass of the hot path over one batch of synthetic LIBERO-shaped observations
(2 x 224px images, prompt of 48 tokens, 64 ActionQuery tokens, proprio, 8x7 chunk; BASELINE.json configs[2]).
One JSON line is printed by rank 0.  `value` is device-resident throughput (inputs already in HBM), `e2e`
goes through the host-buffer C-ABI call (H2D of the observations and D2H of the chunks inside the timed
region), `roofline` describes the dominant kernel (the tcgen05 GEMM) and `cpu_baseline` is the CPU oracle
timed on this box's cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import datetime
import faulthandler
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_T0 = time.perf_counter()
_RANK = int(os.environ.get("RANK", "0"))
# no single stage of the bench takes longer than 240 s; a stage that does is a hang (a profiler that replays every
# kernel needs more: VLA_BENCH_STAGE_LIMIT_S)
STAGE_LIMIT_S = int(os.environ.get("VLA_BENCH_STAGE_LIMIT_S", "240"))

# stdout carries exactly ONE line, the JSON result.  Anything a library prints to fd 1 (NCCL's banner when NCCL_DEBUG is
# set, a stray print) goes to stderr for the whole run; the result line is written to the saved descriptor.
_RESULT_FD = None


def claim_stdout() -> None:
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_result(line: dict) -> None:
    claim_stdout()
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


def stage(msg: str) -> None:
    """Progress of EVERY rank on stderr, and a per-stage dead-man switch: if a rank sits in one stage for more than
    STAGE_LIMIT_S its Python stack is dumped and the process exits non-zero - a hang costs minutes and names its
    stage and rank instead of running into the launcher's limit."""
    print(f"[bench r{_RANK} +{time.perf_counter() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(STAGE_LIMIT_S, exit=True, file=sys.stderr)

METRIC = "action_chunks_per_sec"
UNIT = "chunks/s"
PROMPT_LEN = 48
N_IMAGES = 2
T_CHUNK, A_DIM, P_DIM = 8, 7, 8


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops_sustained": d.get("bf16_tflops_sustained"), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def algorithmic_gflop(n_images: int, L: int, T: int, A: int):
    """Algorithmic FLOPs per sample: what a correct implementation must do.  Excludes the
    discarded last ViT blocks, lm_head, the non-causal half of attention, AQ63/stop rows."""
    def vit(D, Fh, blocks, S):
        lin = S * 2 * (4 * D * D + 2 * D * Fh)
        att = 4 * S * S * D
        return blocks * lin + 2 * 256 * 588 * D, blocks * att
    dl, da = vit(1024, 4096, 23, 261)
    sl, sa = vit(1152, 4304, 26, 256)
    proj = 2 * 256 * (2176 * 8704 + 8704 * 896 + 896 * 896)
    NP = 256 * n_images
    S = NP + L + 63
    llm_lin = 24 * 2 * S * 14909440
    llm_att = 24 * 2 * S * S * 896
    ctx = T + 65 + NP
    pol_lin = 24 * 2 * 896 * 896 * (3 * T + 2 * ctx) + 2 * T * 896 * A
    pol_att = 24 * 4 * T * ctx * 896
    gemm = n_images * (dl + sl + proj) + llm_lin + pol_lin
    att = n_images * (da + sa) + llm_att + pol_att
    return {"total": (gemm + att) / 1e9, "gemm": gemm / 1e9, "attention": att / 1e9}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synth_inputs(B: int, L: int, seed: int, device):
    """Distinct synthetic observations per sample: uint8 images through the processor's two normalisations
    (preprocessor_config.json), random prompt ids, clipped proprio."""
    g = torch.Generator(device=device).manual_seed(1234 + seed)
    img = torch.randint(0, 256, (B, N_IMAGES, 3, 224, 224), generator=g, device=device).float() / 255.0
    m0 = torch.tensor([0.485, 0.456, 0.406], device=device).view(1, 1, 3, 1, 1)
    s0 = torch.tensor([0.229, 0.224, 0.225], device=device).view(1, 1, 3, 1, 1)
    pix = torch.cat([(img - m0) / s0, (img - 0.5) / 0.5], dim=2).reshape(B, 6 * N_IMAGES, 224, 224)
    ids = torch.randint(3, 151643, (B, L), generator=g, device=device, dtype=torch.int64)
    prop = torch.randn(B, P_DIM, generator=g, device=device).clamp(-1, 1)
    return pix.to(torch.bfloat16).contiguous(), ids, prop.float().contiguous()


def workload_name(B: int, L: int, variant: str) -> str:
    return (f"LIBERO predict_action bs={B}/GPU (BASELINE.json configs[2]): 2x224px images, "
            f"L={L} prompt, 64 ActionQuery, proprio, {T_CHUNK}x{A_DIM} chunk, {variant} head")


def gpu_eager_baseline(dev, pro: bool, B: int, L: int):
    """The path's algorithm as plain PyTorch eager ops on the GPU (the oracle's torch code moved to `dev`, bf16 like the
    reference requires): what a user of the reference gets from one B200 today, minus the reference's discarded work
    (lm_head, last ViT blocks).  bs=1 calls (the only mode the reference supports, MP:855) and ONE batched call."""
    from oracle import vla_oracle as O
    ocfg = O.OracleConfig(n_images=N_IMAGES, pro=pro, chunk_len=T_CHUNK, action_dim=A_DIM, proprio_dim=P_DIM,
                          vocab_size=4096)
    OW = {k: v.to(dev, torch.bfloat16) for k, v in O.make_weights(ocfg, seed=0).items()}
    opix, oids, oprop = (t.to(dev) for t in O.make_inputs(ocfg, B, L, seed=0))

    def run(n):
        with torch.no_grad(), torch.device(dev):
            return O.predict_action_batch(OW, ocfg, opix[:n], oids[:n], oprop[:n], torch.bfloat16)["normalized"]

    def timed(n, reps):
        best = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(n)
            torch.cuda.synchronize()
            best.append(time.perf_counter() - t0)
        return statistics.median(best)

    run(1)
    run(1)
    t1 = timed(1, 7)
    run(B)
    tb = timed(B, 3)
    return {"bs1_ms": t1 * 1e3, "bs1_chunks_per_s": 1.0 / t1, "batched_ms": tb * 1e3, "batched_chunks_per_s": B / tb,
            "batch": B, "unit": UNIT,
            "what": "oracle port (torch eager ops: cuBLAS matmuls, softmax(QK^T)V attention) on the same GPU, bf16, "
                    "median wall clock; bs=1 is the reference's only mode, the batched call is what eager PyTorch could do"}


def run_reference(args):
    """--impl reference: the reference's algorithm (oracle port, bf16 like the reference requires) on all host
    cores.  Each step is a bounded sample of the step's batch: ONE observation (1/B of the batch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import vla_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(n_images=N_IMAGES, pro=args.variant == "pro", chunk_len=T_CHUNK, action_dim=A_DIM,
                         proprio_dim=P_DIM, vocab_size=4096)  # the embedding table is only gathered from
    W = O.make_weights(cfg, seed=0)
    pix, ids, prop = O.make_inputs(cfg, 1, PROMPT_LEN, seed=0)
    # The UNMODIFIED reference (its own predict_action through oracle/ref_shim.py) where its sources are mounted -
    # the build container - and the LIBERO constants apply; on the GPU box /root/reference does not exist and the
    # oracle port of the same algorithm is timed instead.
    kind, call = "port", (lambda: O.predict_action_batch(W, cfg, pix, ids, prop, torch.bfloat16))
    try:
        from oracle import ref_shim as R
        if R.available() and (T_CHUNK, A_DIM, P_DIM) == (8, 7, 8):
            stats = {"synthetic": {"action": {"q01": [-1.0] * A_DIM, "q99": [1.0] * A_DIM}}}
            ns, vla, head, pp = R.build_reference(cfg, W, torch.bfloat16, norm_stats=stats)
            kind, call = "reference", (lambda: R.reference_predict_action(ns, vla, head, pp, pix, ids, prop, "synthetic"))
    except Exception as ex:  # the shim needs the reference's sources and transformers; fall back to the port
        print(f"[bench] reference shim unavailable ({type(ex).__name__}: {ex}); timing the oracle port", file=sys.stderr)
    budget_s = args.ref_budget
    times = []
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
        # keep the whole run bounded: once the budget is spent, stop (steps actually timed are reported)
        if time.perf_counter() - t_all > budget_s and len(times) >= 1:
            break
    ms = statistics.mean(times) * 1e3
    val = 1000.0 / ms
    sample = (f"1 observation per step (1/{args.batch} of the bs={args.batch} batch), bf16 on the host CPU, "
              + ("the unmodified reference's predict_action" if kind == "reference" else "torch-CPU oracle port")
              + f", {len(times)} timed steps of the requested {args.steps}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        # the same workload as our arm; this arm times a bounded sample of it (one observation per step, see `sample`)
        "config": {"workload": workload_name(args.batch, PROMPT_LEN, args.variant), "global_batch": args.gpus * args.batch,
                   "per_gpu_batch": args.batch, "sampled_observations_per_step": 1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_result(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU per step")
    ap.add_argument("--variant", default="base", choices=["base", "pro"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=200.0, help="wall-clock bound [s] of the reference arm")
    ap.add_argument("--latency-iters", type=int, default=200, help="bs=1 latency samples (after 20 warm-up calls)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch samples per GPU; strong: --batch is the GLOBAL batch, split over the ranks "
                         "(SURVEY 8d config 4 asks for both)")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--chunk", default="8x7x8", help="chunk_len x action_dim x proprio_dim: 8x7x8 = LIBERO / CALVIN "
                    "(constants.py:28-40), 25x14x14 = the reference's larger-chunk preset (ALOHA, constants.py:42-47)")
    args = ap.parse_args()
    claim_stdout()
    global T_CHUNK, A_DIM, P_DIM
    T_CHUNK, A_DIM, P_DIM = (int(v) for v in args.chunk.split("x"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        stage(f"init_process_group(nccl), world {world}")
        # a collective that does not complete within two minutes is a failure, not something to wait 10 minutes for
        # (the reference sets an explicit process-group timeout too, vla-scripts/evaluate_calvin.py:877); the flight
        # recorder then says which collective of which rank was outstanding
        os.environ.setdefault("TORCH_NCCL_TRACE_BUFFER_SIZE", "2000")
        os.environ.setdefault("TORCH_NCCL_DUMP_ON_TIMEOUT", "1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
        dist.barrier()
        torch.cuda.synchronize()
        stage("process group up")

    from vla_adapter_b200 import _lib
    from vla_adapter_b200.engine import VLAEngine
    from vla_adapter_b200.weights import load_random_weights
    from vla_adapter_b200 import tokens
    from vla_adapter_b200 import sharding

    B, L, K, Wu = args.batch, PROMPT_LEN, args.steps, args.warmup
    if args.scaling == "strong":
        if B % world:
            raise SystemExit(f"--scaling strong: the global batch {B} must divide over {world} ranks")
        B //= world
    pro = args.variant == "pro"
    eng = VLAEngine(n_images=N_IMAGES, chunk_len=T_CHUNK, action_dim=A_DIM, proprio_dim=P_DIM, pro=pro, max_batch=B,
                    max_prompt_len=L, device=local,
                    norm_stats={"synthetic": {"action": {"q01": [-1.0] * A_DIM, "q99": [1.0] * A_DIM,
                                                         "mask": [True] * (A_DIM - 1) + [False]}}})
    n_params = load_random_weights(eng, seed=0, n_images=N_IMAGES, action_dim=A_DIM, proprio_dim=P_DIM, pro=pro)
    eng.finalize()
    stage("engine built: random weights loaded, finalized")
    lib = _lib.load()

    pix, ids, prop = synth_inputs(B, L, seed=rank, device=dev)
    ext, _, _, aq, _ = tokens.build(ids.cpu(), None, A_DIM)
    ext_d, aq_d = ext.to(dev), aq.to(dev)

    def step_device():
        out_n, out_u, _ = eng.predict_device(pix, ext_d, aq_d, prop)
        if world > 1:
            # the only collective of the path: the product's own gather of action chunks over NVLink
            sharding.gather_chunks(out_u, world * B)
        return out_u

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput
    stage("warm-up + timed steps (device-resident inputs)")
    for _ in range(Wu):
        step_device()
    sync_all()
    launches0 = lib.vla_total_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(K):
            out = step_device()
        e1.record()
        sync_all()
    ms_step = reduce_max(e0.elapsed_time(e1) / K)
    # kernels launched inside the timed region: the engine replays a captured CUDA graph of its forward, so the
    # per-step count is the one recorded at capture time (vla_last_launch_count), not a host-side counter delta
    launches = max(lib.vla_total_launch_count() - launches0, eng.last_launch_count() * K)
    value = world * B / (ms_step * 1e-3)
    assert torch.isfinite(out).all(), "non-finite action chunk"

    # ---------------- end to end through the host-buffer C-ABI call (pinned host memory)
    stage("end-to-end steps through the host-buffer call")
    pix_h, ext_h, aq_h, prop_h = pix.cpu().pin_memory(), ext.pin_memory(), aq.pin_memory(), prop.cpu().pin_memory()
    on_h = torch.empty((B, T_CHUNK, A_DIM), dtype=torch.float32).pin_memory()
    ou_h = torch.empty((B, T_CHUNK, A_DIM), dtype=torch.float32).pin_memory()
    for _ in range(2):
        eng.predict_host(pix_h, ext_h, aq_h, prop_h, on_h, ou_h)
    sync_all()
    e0.record()
    for _ in range(K):
        eng.predict_host(pix_h, ext_h, aq_h, prop_h, on_h, ou_h)
        if world > 1:
            sharding.gather_chunks(ou_h.to(dev, non_blocking=True), world * B)
    e1.record()
    sync_all()
    ms_e2e = reduce_max(e0.elapsed_time(e1) / K)
    h2d = pix_h.numel() * 2 + ext_h.numel() * 8 + aq_h.numel() * 4 + prop_h.numel() * 4
    d2h = 2 * on_h.numel() * 4 + 4
    assert torch.equal(ou_h, out.cpu()), "host-path result differs from device-path result"

    # the same end-to-end call from uint8 frames (device-side ToTensor + Normalize): 4x fewer H2D bytes
    g8 = torch.Generator().manual_seed(99 + rank)
    img_h = torch.randint(0, 256, (B, N_IMAGES, 224, 224, 3), generator=g8, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        eng.predict_host_u8(img_h, ext_h, aq_h, prop_h, on_h, ou_h)
    sync_all()
    e0.record()
    for _ in range(K):
        eng.predict_host_u8(img_h, ext_h, aq_h, prop_h, on_h, ou_h)
    e1.record()
    sync_all()
    ms_e2e_u8 = reduce_max(e0.elapsed_time(e1) / K)
    h2d_u8 = img_h.numel() + ext_h.numel() * 8 + aq_h.numel() * 4 + prop_h.numel() * 4

    # ---------------- roofline of the dominant kernel: tcgen05 GEMM, timed per launch with CUDA events
    stage("per-launch GEMM timing + segment timing")
    lib.vla_profile_gemm(1)
    step_device()
    torch.cuda.synchronize()
    import ctypes as C
    g_ms, g_n = C.c_double(0), C.c_longlong(0)
    lib.vla_profile_gemm_read(C.byref(g_ms), C.byref(g_n))
    lib.vla_profile_gemm(0)
    peaks = measured_peaks()
    alg = algorithmic_gflop(N_IMAGES, L, T_CHUNK, A_DIM)
    gemm_tflops = alg["gemm"] * B / g_ms.value  # GFLOP / ms = TFLOP/s
    peak = peaks["tflops_sustained"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": gemm_tflops, "peak": peak,
                "unit": "TFLOP/s", "frac": gemm_tflops / peak, "traffic": traffic,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": int(g_n.value), "avg_launch_ms": g_ms.value / max(1, g_n.value),
                "gemm_ms_per_step": g_ms.value, "gemm_share_of_step": g_ms.value / ms_step,
                "algorithmic_gflop_per_sample": alg}
    # per-subsystem device time of one eager step (events at the subsystem boundaries inside the engine)
    seg3 = (C.c_float * 3)()
    lib.vla_segment_timing(eng._h, 1)
    step_device()
    step_device()
    torch.cuda.synchronize()
    _lib.check(lib.vla_segment_times(eng._h, seg3), eng._h)
    lib.vla_segment_timing(eng._h, 0)
    NPt = 256 * N_IMAGES
    fused_gf = alg["total"] - N_IMAGES * 382.77  # prefill + policy = total minus the towers/projector
    segments = {"towers_projector_ms": seg3[0], "llm_prefill_ms": seg3[1], "policy_ms": seg3[2],
                "prefill_policy": {"gflop_per_sample": fused_gf, "achieved": fused_gf * B / (seg3[1] + seg3[2]),
                                   "peak": peak, "unit": "TFLOP/s",
                                   "frac": fused_gf * B / (seg3[1] + seg3[2]) / peak,
                                   "note": "north_star sub-path (Qwen prefill + Bridge-Attention policy), eager step"},
                "towers_projector": {"achieved": N_IMAGES * 382.77 * B / seg3[0], "unit": "TFLOP/s",
                                     "frac": N_IMAGES * 382.77 * B / seg3[0] / peak}}
    step_tflops = alg["total"] * B / ms_step
    step_roofline = {"achieved": step_tflops, "peak": peak, "unit": "TFLOP/s", "frac": step_tflops / peak,
                     "note": "whole step (all kernels) against the same measured dense-bf16 peak"}

    # ---------------- bs=1 latency through the host path (p50 / p90), BASELINE.json's second metric
    stage("bs=1 latency")
    one = [t[:1].contiguous().pin_memory() for t in (pix_h, ext_h, aq_h, prop_h)]
    o1, o2 = on_h[:1].clone().pin_memory(), ou_h[:1].clone().pin_memory()

    def bs1_latency(engine, head):
        lat = []
        for i in range(args.latency_iters + 20 if args.latency_iters > 0 else 0):
            t0 = time.perf_counter()
            engine.predict_host(one[0], one[1], one[2], one[3], o1, o2)
            if i >= 20:
                lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        if not lat:
            return None
        return {"head": head, "p50_ms": lat[len(lat) // 2], "p90_ms": lat[int(len(lat) * 0.9)], "iters": len(lat),
                "warmup": 20, "how": "wall clock around vla_predict_host(B=1) incl. H2D/D2H and stream sync"}

    latency = bs1_latency(eng, args.variant)
    # BASELINE.json configs[1] asks for the bs=1 latency of BOTH heads: a second, bs=1-sized engine with the other head
    latency_other = None
    if world == 1 and latency is not None:
        other = "base" if pro else "pro"
        stage(f"bs=1 latency, {other} head")
        eng1 = VLAEngine(n_images=N_IMAGES, chunk_len=T_CHUNK, action_dim=A_DIM, proprio_dim=P_DIM, pro=not pro,
                         max_batch=1, max_prompt_len=L, device=local)
        load_random_weights(eng1, seed=0, n_images=N_IMAGES, action_dim=A_DIM, proprio_dim=P_DIM, pro=not pro)
        eng1.finalize()
        latency_other = bs1_latency(eng1, other)
        eng1.close()

    # ---------------- CPU baseline (rank 0, N=1 only): the oracle on this box's cores, bounded sample
    stage("CPU baseline (oracle)")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import vla_oracle as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ocfg = O.OracleConfig(n_images=N_IMAGES, pro=pro, chunk_len=T_CHUNK, action_dim=A_DIM, proprio_dim=P_DIM,
                              vocab_size=4096)
        OW = O.make_weights(ocfg, seed=0)
        opix, oids, oprop = O.make_inputs(ocfg, 1, L, seed=0)
        O.predict_action_batch(OW, ocfg, opix, oids, oprop, torch.bfloat16)  # warm-up (oneDNN primitive caches)
        dts = []
        t_begin = time.perf_counter()
        while len(dts) < 5 and (time.perf_counter() - t_begin < 20.0 or not dts):
            t0 = time.perf_counter()
            O.predict_action_batch(OW, ocfg, opix, oids, oprop, torch.bfloat16)
            dts.append(time.perf_counter() - t0)
        dt = statistics.median(dts)
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "1 observation per call (1/%d of the batch), bf16 torch-CPU oracle, median of %d calls after "
                         "1 warm-up, embedding table cut to 4096 rows (gather only)" % (B, len(dts))}

    # ---------------- like-for-like GPU baseline: the same algorithm in PyTorch eager on this GPU
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager_baseline:
        stage("GPU eager baseline (oracle port in PyTorch eager, bf16, on this GPU)")
        try:
            gpu_eager = gpu_eager_baseline(dev, pro, B, L)
        except Exception as ex:  # a baseline must never take the measurement down with it
            gpu_eager = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wu,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(B, L, args.variant),
                       "global_batch": world * B, "per_gpu_batch": B, "params": n_params,
                       "parallelism": f"sample-sharded x{world}, full weight replica per GPU, all-gather of chunks",
                       "l2": "no explicit flush: per-step working set (2.7 GB weights + >4 GB activations) >> 126 MB L2"},
            "clocks": clk.summary(),
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_uint8_frames": {"value": world * B / (ms_e2e_u8 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_u8,
                                 "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                                 "note": "vla_predict_host_u8: uint8 HWC frames, normalisation on the device"},
            "gpu_launches": int(launches),
            "roofline": roofline, "step_roofline": step_roofline, "segments": segments, "latency_bs1": latency,
            "latency_bs1_other_head": latency_other, "gpu_eager_baseline": gpu_eager,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit_result(line)
    stage("done")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    main()
