"""Oracle parity on the code path the headline number runs: B > 8 takes the engine's LARGE-BATCH branch (stand-alone
RoPE kernel, one shared policy K|V buffer, no side stream, no programmatic dependent launch), which the B <= 3 cases
of test_engine_gpu.py never reach.  Also SURVEY 8d's gate on >= 32 seeded samples per head variant, the three-image
input (S = 881, deploy.py:128's default), the asynchronous error contract of the device-pointer call, and two engines
alive in one process."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vla_oracle as O  # noqa: E402  (the checker)

TAPS = ["patches", "projected", "llm_in", "hidden.1", "hidden.12", "hidden.24", "head_x.0", "head_x.1", "head_x.12",
        "head_x.24"]


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _engine(cfg, W, B, L, **kw):
    from vla_adapter_b200.engine import VLAEngine

    eng = VLAEngine(n_images=cfg.n_images, chunk_len=cfg.chunk_len, action_dim=cfg.action_dim,
                    proprio_dim=cfg.proprio_dim, pro=cfg.pro, dino_depth=cfg.dino_depth, siglip_depth=cfg.siglip_depth,
                    vocab_size=cfg.vocab_size, max_batch=B, max_prompt_len=L, **kw)
    eng.load_flat(W)
    eng.finalize()
    return eng


@pytest.mark.parametrize("pro", [False, True])
def test_large_batch_branch_gate_32_samples(pro):
    """2 seeds x 16 samples per variant (>= 32) through max_batch = 16 > 8: taps and actions against the
    fp32 truth, held to twice the reference-precision (bf16) oracle's own error."""
    cfg = O.OracleConfig(n_images=2, dino_depth=3, siglip_depth=3, vocab_size=2048, pro=pro)
    W = O.make_weights(cfg, seed=11)
    B, L = 16, 29
    eng = _engine(cfg, W, B, L)
    worst = 0.0
    for seed in (21, 22):
        pix, ids, prop = O.make_inputs(cfg, B, L, seed=seed)
        truth = O.predict_action_batch(W, cfg, pix, ids, prop, torch.float32, keep_taps=True)
        ref16 = O.predict_action_batch(W, cfg, pix, ids, prop, torch.bfloat16, keep_taps=True)
        actions, normalized, ha = eng.predict_action_batch(ids, None, pix, prop, return_hidden=True)
        for k in TAPS:
            t = truth[k].float().reshape(-1)
            r, g = _rel(ref16[k].reshape(-1), t), _rel(eng.tap(k).float().cpu(), t)
            assert g <= max(2 * r, 1e-2), f"{k}: engine {g:.4f} vs bf16 reference {r:.4f} (seed {seed})"
        tn = truth["normalized"].numpy()
        e_ref, e_eng = np.abs(ref16["normalized"].numpy() - tn), np.abs(normalized - tn)
        print(f"pro={pro} seed={seed}: engine max {e_eng.max():.4f} mean {e_eng.mean():.4f} | "
              f"bf16 oracle max {e_ref.max():.4f} mean {e_ref.mean():.4f}")
        assert e_eng.max() <= max(2 * e_ref.max(), 2e-2)
        assert e_eng.mean() <= 5e-3 + e_ref.mean()
        worst = max(worst, e_eng.max())
        assert _rel(ha.float().reshape(-1), truth["last_ha"].float().reshape(-1)) <= max(
            2 * _rel(ref16["last_ha"].reshape(-1), truth["last_ha"].reshape(-1)), 1e-2)
    # the same engine, same samples through the SMALL-batch branch (B = 4 <= 8): both branches agree to bf16 noise
    pix, ids, prop = O.make_inputs(cfg, B, L, seed=22)
    _, n_small = eng.predict_action_batch(ids[:4], None, pix[:4], prop[:4])
    _, n_large = eng.predict_action_batch(ids, None, pix, prop)
    eng.close()
    # small batches run the policy as one cluster kernel, large ones per block: same roundings, another summation order;
    # the Pro head (RoPE + gated task/adapter mix) amplifies that more than the base head
    assert np.abs(n_small - n_large[:4]).max() <= (6e-2 if pro else 2e-2)


def test_three_images():
    """n_images = 3 (deploy.py:128's default): S = 768 + L + 65; the causal LLM attention has 7 x 7 = 49 work units per
    kv head and must stay on the tcgen05 kernel.  (The reference's head hard-wires num_task_tokens = 512, AH:28; the
    engine and the oracle split at NP = 768 - an intentional deviation like the one-image case.)"""
    cfg = O.OracleConfig(n_images=3, dino_depth=2, siglip_depth=2, vocab_size=1024, pro=False)
    W = O.make_weights(cfg, seed=12)
    B, L = 2, 48
    pix, ids, prop = O.make_inputs(cfg, B, L, seed=12)
    truth = O.predict_action_batch(W, cfg, pix, ids, prop, torch.float32, keep_taps=True)
    ref16 = O.predict_action_batch(W, cfg, pix, ids, prop, torch.bfloat16, keep_taps=True)
    from vla_adapter_b200 import _lib

    lib = _lib.load()
    lib.vla_set_attention_impl(2)  # tcgen05 or fail: no silent drop to the mma.sync kernel
    try:
        eng = _engine(cfg, W, B, L)
        _, normalized = eng.predict_action_batch(ids, None, pix, prop)
        for k in ("hidden.1", "hidden.24", "head_x.24"):
            t = truth[k].float().reshape(-1)
            assert _rel(eng.tap(k).float().cpu(), t) <= max(2 * _rel(ref16[k].reshape(-1), t), 1e-2), k
        eng.close()
    finally:
        lib.vla_set_attention_impl(0)
    tn = truth["normalized"].numpy()
    assert np.abs(normalized - tn).max() <= max(2 * np.abs(ref16["normalized"].numpy() - tn).max(), 2e-2)


def test_device_call_reports_bad_ids_through_check_errors():
    """vla_predict only enqueues; vla_check_errors reports what the kernels found and clears it.  The offending row
    is zero-filled (never a stale row of an earlier call), so the bad call's result is deterministic."""
    cfg = O.OracleConfig(n_images=1, dino_depth=2, siglip_depth=2, vocab_size=512, pro=False)
    W = O.make_weights(cfg, seed=13)
    B, L = 2, 9
    eng = _engine(cfg, W, B, L)
    pix, ids, prop = O.make_inputs(cfg, B, L, seed=13)
    dev = eng.device
    pix_d, prop_d = pix.to(dev).to(torch.bfloat16).contiguous(), prop.to(dev).float().contiguous()

    def run(ids_cpu):
        ext, aq, _ = eng._prep(ids_cpu, None)
        out = eng.predict_device(pix_d, ext.to(dev), aq.to(dev), prop_d)[0]
        rc = eng.lib.vla_check_errors(eng._h, torch.cuda.current_stream().cuda_stream)
        return out.cpu(), rc

    good, rc = run(ids)
    assert rc == 0
    bad = ids.clone()
    bad[1, 4] = 512
    out1, rc = run(bad)
    assert rc == -1 and b"token id" in eng.lib.vla_last_error(eng._h)
    other = ids.clone()
    other[1] = (ids[1] + 7) % 512
    run(other)                      # a different valid call in between leaves different rows behind
    out2, rc2 = run(bad)
    assert rc2 == -1
    assert torch.equal(out1, out2), "a rejected id must not expose rows of an earlier call"
    assert torch.equal(out1[0], good[0])  # the other sample of the batch is untouched
    again, rc = run(ids)
    eng.close()
    assert rc == 0 and torch.equal(again, good)


def test_two_engines_interleaved():
    """Two engines in one process (different heads and shapes), calls interleaved: per-engine state only."""
    cfg_a = O.OracleConfig(n_images=1, dino_depth=2, siglip_depth=2, vocab_size=512, pro=False)
    cfg_b = O.OracleConfig(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=512, pro=True)
    Wa, Wb = O.make_weights(cfg_a, seed=14), O.make_weights(cfg_b, seed=15)
    ea, eb = _engine(cfg_a, Wa, 2, 10), _engine(cfg_b, Wb, 12, 10)
    ia, ib = O.make_inputs(cfg_a, 2, 10, seed=1), O.make_inputs(cfg_b, 12, 10, seed=2)
    ra = [ea.predict_action_batch(ia[1], None, ia[0], ia[2])[1]]
    rb = [eb.predict_action_batch(ib[1], None, ib[0], ib[2])[1]]
    for _ in range(2):
        ra.append(ea.predict_action_batch(ia[1], None, ia[0], ia[2])[1])
        rb.append(eb.predict_action_batch(ib[1], None, ib[0], ib[2])[1])
    ea.close()
    rb.append(eb.predict_action_batch(ib[1], None, ib[0], ib[2])[1])
    eb.close()
    assert all(np.array_equal(r, ra[0]) for r in ra) and all(np.array_equal(r, rb[0]) for r in rb)
    ta = O.predict_action_batch(Wa, cfg_a, *ia, torch.float32)["normalized"].numpy()
    tb = O.predict_action_batch(Wb, cfg_b, *ib, torch.float32)["normalized"].numpy()
    assert np.abs(ra[0] - ta).max() < 5e-2 and np.abs(rb[0] - tb).max() < 5e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_one_process():
    """per-device kernel attributes / SM counts and a device guard at every entry point - an engine on
    cuda:1 works after one on cuda:0 was built, whatever device is current at call time."""
    cfg = O.OracleConfig(n_images=1, dino_depth=2, siglip_depth=2, vocab_size=512, pro=False)
    W = O.make_weights(cfg, seed=16)
    pix, ids, prop = O.make_inputs(cfg, 2, 10, seed=3)
    e0 = _engine(cfg, W, 2, 10, device=0)
    e1 = _engine(cfg, W, 2, 10, device=1)
    torch.cuda.set_device(0)
    n1 = e1.predict_action_batch(ids, None, pix, prop)[1]   # current device 0, engine on 1
    torch.cuda.set_device(1)
    n0 = e0.predict_action_batch(ids, None, pix, prop)[1]   # and the other way round
    torch.cuda.set_device(0)
    e0.close()
    e1.close()
    assert np.array_equal(n0, n1)


@pytest.mark.parametrize("pro,B", [(False, 5), (True, 12)])
def test_mixed_prompt_lengths_in_one_batch(pro, B, monkeypatch):
    """Per-sample prompt lengths (vla_predict's prompt_len; chat prompts span 40-56 tokens, OU:783): the batch is
    right-padded, causal attention hides the padding, and every sample must come out exactly as when it runs alone
    with its own length - and within the usual gate of the oracle run per sample (the reference is bs=1, MP:855)."""
    cfg = O.OracleConfig(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=1024, pro=pro)
    W = O.make_weights(cfg, seed=17)
    lens = [9, 14, 7, 14, 11, 8, 13, 14, 10, 12, 9, 14][:B]
    Lmax = max(lens)
    pix, ids, prop = O.make_inputs(cfg, B, Lmax, seed=17)
    mask = (torch.arange(Lmax)[None] < torch.tensor(lens)[:, None]).long()
    if B > 8:
        # the stand-alone runs (B = 1) would take the single-kernel policy (policy_fused.cu), the batch the per-block
        # tcgen05 chain: same roundings, different fp32 summation order - and the Pro head turns that into ~2e-2 on the
        # actions.  This test is about padding, so both sides use the per-block chain; the two policy paths are
        # compared in test_fused_policy_matches_per_block_chain.
        monkeypatch.setenv("VLA_NO_POLICY_FUSED", "1")
    eng = _engine(cfg, W, B, Lmax)
    actions, normalized, ha = eng.predict_action_batch(ids, mask, pix, prop, return_hidden=True)
    # list-of-rows form of the same call
    _, n_list = eng.predict_action_batch([ids[b, :lens[b]] for b in range(B)], None, pix, prop)
    assert np.array_equal(n_list, normalized)
    worst_alone, worst_oracle, worst_ref = 0.0, 0.0, 0.0
    for b in range(B):
        i1 = ids[b:b + 1, :lens[b]]
        _, n1, h1 = eng.predict_action_batch(i1, None, pix[b:b + 1], prop[b:b + 1], return_hidden=True)
        worst_alone = max(worst_alone, float(np.abs(n1[0] - normalized[b]).max()))
        assert torch.equal(h1[0], ha[b]), f"sample {b}: last-layer ActionQuery states differ from the stand-alone run"
        if b < 5:   # the oracle (fp32 truth and the reference-precision bf16 run) on its own un-padded prompt
            t = O.predict_action_batch(W, cfg, pix[b:b + 1], i1, prop[b:b + 1], torch.float32)["normalized"].numpy()
            r = O.predict_action_batch(W, cfg, pix[b:b + 1], i1, prop[b:b + 1], torch.bfloat16)["normalized"].numpy()
            worst_oracle = max(worst_oracle, float(np.abs(t[0] - normalized[b]).max()))
            worst_ref = max(worst_ref, float(np.abs(t[0] - r[0]).max()))
    eng.close()
    print(f"pro={pro} B={B}: max |batched - alone| = {worst_alone:.2e}, max |batched - fp32 oracle| = {worst_oracle:.4f} "
          f"(bf16 oracle {worst_ref:.4f})")
    assert worst_alone <= 2e-3       # same arithmetic per row; only the batch-size branch (B = 1 vs B) may differ
    assert worst_oracle <= max(2 * worst_ref, 2e-2)
    with pytest.raises(ValueError):  # a zero-length prompt
        eng2 = _engine(cfg, W, 2, 4)
        try:
            eng2.predict_action_batch(ids[:2, :4], torch.tensor([[1, 1, 1, 1], [0, 0, 0, 0]]), pix[:2], prop[:2])
        finally:
            eng2.close()


def test_action_batcher_on_a_real_engine():
    """SURVEY 8f-3 end to end: concurrent /act-style requests with different prompt lengths share ONE forward of a
    real engine, and every caller gets the chunk the engine gives that observation alone."""
    import threading

    from vla_adapter_b200.serving import ActionBatcher

    cfg = O.OracleConfig(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=1024, pro=False)
    W = O.make_weights(cfg, seed=18)
    stats = {"k": {"action": {"q01": [-0.5] * 7, "q99": [0.75] * 7, "mask": [True] * 6 + [False]}}}
    lens = [40, 56, 48, 44, 52, 41]
    eng = _engine(cfg, W, len(lens), 56, norm_stats=stats)
    pix, ids, prop = O.make_inputs(cfg, len(lens), 56, seed=18)
    alone = [eng.predict_action_batch(ids[b:b + 1, :lens[b]], None, pix[b:b + 1], prop[b:b + 1], unnorm_key="k")[0][0]
             for b in range(len(lens))]
    batcher = ActionBatcher(eng, max_batch=len(lens), max_wait_ms=500)
    got, errs = {}, {}

    def go(b):
        try:
            got[b] = batcher.submit(ids[b, :lens[b]], pix[b], prop[b].numpy(), "k", timeout=60)
        except Exception as ex:  # noqa: BLE001
            errs[b] = ex

    ts = [threading.Thread(target=go, args=(b,)) for b in range(len(lens))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(90)
    batcher.close()
    eng.close()
    assert not errs, errs
    assert sum(batcher.batches) == len(lens) and max(batcher.batches) >= 2, batcher.batches
    for b in range(len(lens)):
        assert got[b].shape == (8, 7) and got[b].dtype == np.float64
        assert np.abs(got[b] - alone[b]).max() <= 2e-3, b


@pytest.mark.parametrize("pro", [False, True])
def test_fused_policy_matches_per_block_chain(pro, monkeypatch):
    """Small batches run the 24 policy blocks as one cluster kernel (policy_fused.cu, mma.sync, weights streamed from
    L2); VLA_NO_POLICY_FUSED=1 keeps the per-block launches (tcgen05 GEMMs + split-kv attention).  Both keep every bf16
    rounding of AH:218-283 / 337-410 at the same place, so the final policy state agrees to fp32 summation order and
    both pass the oracle gate."""
    cfg = O.OracleConfig(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=1024, pro=pro)
    W = O.make_weights(cfg, seed=23)
    B, L = 3, 13
    pix, ids, prop = O.make_inputs(cfg, B, L, seed=23)
    truth = O.predict_action_batch(W, cfg, pix, ids, prop, torch.float32, keep_taps=True)
    ref16 = O.predict_action_batch(W, cfg, pix, ids, prop, torch.bfloat16, keep_taps=True)
    out = {}
    for fused in (True, False):
        if fused:
            monkeypatch.delenv("VLA_NO_POLICY_FUSED", raising=False)
        else:
            monkeypatch.setenv("VLA_NO_POLICY_FUSED", "1")
        eng = _engine(cfg, W, B, L)
        _, n = eng.predict_action_batch(ids, None, pix, prop)
        out[fused] = (n, eng.tap("head_x.24").float().cpu().reshape(-1), eng.tap("head_x.1").float().cpu().reshape(-1))
        eng.close()
    t24 = truth["head_x.24"].float().reshape(-1)
    for fused in (True, False):
        assert _rel(out[fused][1], t24) <= max(2 * _rel(ref16["head_x.24"].reshape(-1), t24), 1e-2), f"fused={fused}"
    assert _rel(out[True][2], out[False][2]) <= 4e-3      # after one block: bf16 rounding noise only
    assert _rel(out[True][1], out[False][1]) <= 2e-2      # after 24
    tn = truth["normalized"].numpy()
    gate = max(2 * np.abs(ref16["normalized"].numpy() - tn).max(), 2e-2)
    assert np.abs(out[True][0] - tn).max() <= gate and np.abs(out[False][0] - tn).max() <= gate
