"""Full-size checks (BASELINE.json configs[2]: full-depth towers / LLM / policy, bs=64, 2 images, L=48) through
properties that need no oracle run - the CPU oracle takes minutes per sample at this size:

  * sample permutation: the path has no cross-sample interaction, so permuting the batch permutes the
    outputs BIT-EXACTLY (each row's arithmetic does not depend on the tile it lands in);
  * sub-batch: the first 8 samples alone (small-batch mode: side stream, PDL, fused RoPE, other tile widths) agree
    with their rows of the bs=64 run;
  * base head: all T rows of a chunk are identical; Pro head (RoPE): they differ;
  * un-normalisation: out_unnorm = 0.5 (a + 1)(q99 - q01 + 1e-8) + q01 on masked dims, a elsewhere (MP:786-805);
  * replay: the captured CUDA graph reproduces the eager result.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B, L, device, seed=5):
    g = torch.Generator(device=device).manual_seed(seed)
    img = torch.randint(0, 256, (B, 2, 3, 224, 224), generator=g, device=device).float() / 255.0
    m0 = torch.tensor([0.485, 0.456, 0.406], device=device).view(1, 1, 3, 1, 1)
    s0 = torch.tensor([0.229, 0.224, 0.225], device=device).view(1, 1, 3, 1, 1)
    pix = torch.cat([(img - m0) / s0, (img - 0.5) / 0.5], dim=2).reshape(B, 12, 224, 224)
    ids = torch.randint(3, 151643, (B, L), generator=g, device=device, dtype=torch.int64)
    prop = torch.randn(B, 8, generator=g, device=device).clamp(-1, 1)
    return pix.to(torch.bfloat16).contiguous(), ids, prop.float().contiguous()


@pytest.mark.parametrize("pro", [False, True])
def test_full_size_batch_properties(pro):
    from vla_adapter_b200 import tokens
    from vla_adapter_b200.engine import VLAEngine
    from vla_adapter_b200.weights import load_random_weights

    B, L = 64, 48
    q01 = [-1.0, -0.5, -2.0, -1.0, -0.25, -3.0, 0.0]
    q99 = [1.0, 0.5, 1.0, 3.0, 0.25, 0.0, 1.0]
    mask = [True] * 6 + [False]
    eng = VLAEngine(n_images=2, pro=pro, max_batch=B, max_prompt_len=L,
                    norm_stats={"synthetic": {"action": {"q01": q01, "q99": q99, "mask": mask}}})
    try:
        load_random_weights(eng, seed=0, n_images=2, action_dim=7, proprio_dim=8, pro=pro)
        eng.finalize()
        dev = torch.device("cuda", torch.cuda.current_device())
        pix, ids, prop = _inputs(B, L, dev)
        ext, _, _, aq, _ = tokens.build(ids.cpu(), None, 7)
        ext_d, aq_d = ext.to(dev), aq.to(dev)

        def run(sel=None):
            a = (pix, ext_d, aq_d, prop) if sel is None else tuple(t[sel].contiguous() for t in (pix, ext_d, aq_d, prop))
            n, u, h = eng.predict_device(*a, want_last_ha=True)
            torch.cuda.synchronize()
            return n.clone(), u.clone(), h.clone()

        n1, u1, h1 = run()            # eager
        n1b, u1b, h1b = run()         # captured on this call
        n1c, _, _ = run()             # graph replay
        assert torch.isfinite(n1).all() and torch.isfinite(h1.float()).all()
        assert torch.equal(n1, n1b) and torch.equal(n1, n1c) and torch.equal(h1, h1b)
        assert n1.std().item() > 1e-3 and (n1[0] - n1[1]).abs().max().item() > 0   # samples really differ

        perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).to(dev)
        n2, u2, h2 = run(perm)
        assert torch.equal(n2, n1[perm]) and torch.equal(u2, u1[perm]) and torch.equal(h2, h1[perm])

        n3, u3, h3 = run(torch.arange(8, device=dev))
        assert (n3 - n1[:8]).abs().max().item() <= (6e-2 if pro else 2e-2)   # other policy path at B <= 8 (policy_fused.cu)
        assert (h3.float() - h1[:8].float()).abs().max().item() <= 2e-2 * max(1.0, h1.float().abs().max().item())

        if not pro:      # no positional signal in the base head; the Pro head's RoPE tells the rows apart
            for t in range(1, 8):
                assert torch.equal(n1[:, 0], n1[:, t])
        else:
            assert not torch.equal(n1[:, 0], n1[:, 1])

        a = n1.cpu().numpy().astype(np.float64)
        lo, hi = np.asarray(q01), np.asarray(q99)
        want = np.where(np.asarray(mask), 0.5 * (a + 1.0) * (hi - lo + 1e-8) + lo, a)
        assert np.allclose(u1.cpu().numpy().astype(np.float64), want, rtol=1e-5, atol=1e-6)
    finally:
        eng.close()
