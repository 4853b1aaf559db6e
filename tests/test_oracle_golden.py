"""Pins the CPU oracle (oracle/vla_oracle.py) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz; the reference ships no tests or vectors of its own).

  integer path          bit-exact (np.array_equal) against MP:748-784 / TU:8-41 outputs
  vision + projector    fp32 oracle == reference fp32 sub-modules, rel-L2 <= 1e-5
  Qwen2.5 prefill       fp32 oracle vs HF Qwen2 fp32 hidden states 1/12/24, rel-L2 <= 1e-5
  policy head           bf16 oracle head == reference L1RegressionActionHead on the same input, BIT-EXACT
                        (both variants; this is what caught the deployed bf16 inv_freq of the Pro RoPE)
  end to end            un-normalised actions: |oracle_fp32 - reference_bf16| <= 4e-2 (bf16 path noise on
                        [-1, 1]-scale outputs), last-layer ActionQuery states rel-L2 <= 3e-2
"""
import os

import numpy as np
import pytest
import torch

from oracle import vla_oracle as O
from oracle.make_golden import CASES, STATS, case_config, checksum, weights_checksum

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    name = request.param
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg, B, L, seed = case_config(name)
    W = O.make_weights(cfg, seed=seed)
    pix, ids, prop = O.make_inputs(cfg, B, L, seed=seed)
    # the seeded problem must be the one the golden file was generated from
    assert weights_checksum(W) == int(g["w_crc"]), "seeded weights drifted from the golden file"
    assert checksum(pix) == int(g["pix_crc"]), "seeded images drifted from the golden file"
    assert np.array_equal(ids.numpy(), g["ids"])
    o32 = O.predict_action_batch(W, cfg, pix, ids, prop, torch.float32, keep_taps=True)
    o16 = O.predict_action_batch(W, cfg, pix, ids, prop, torch.bfloat16, keep_taps=True)
    return name, g, cfg, W, (pix, ids, prop), o32, o16


def test_integer_path_bit_exact(case):
    name, g, cfg, W, (pix, ids, prop), o32, _ = case
    assert np.array_equal(o32["ext_ids"].numpy(), g["ref_ext_ids"])
    assert np.array_equal(o32["labels"].numpy(), g["ref_labels"])
    assert np.array_equal(o32["mask"].numpy(), g["ref_mask"])
    L = ids.shape[1]
    assert g["ref_mask"][:, L:L + 64].all() and g["ref_mask"].sum() == 64 * ids.shape[0]
    # the product's host logic (vla_adapter_b200/tokens.py) gives the same integers
    from vla_adapter_b200 import tokens

    ext, labels, mask, aq, ext_mask = tokens.build(ids, torch.ones_like(ids), cfg.action_dim)
    assert np.array_equal(ext.numpy(), g["ref_ext_ids"])
    assert np.array_equal(labels.numpy(), g["ref_labels"])
    assert np.array_equal(mask.numpy(), g["ref_mask"])
    assert np.array_equal(ext_mask.numpy(), g["ref_attention_mask"])
    assert np.array_equal(aq.numpy(), o32["aq_index"].numpy())


def test_fp32_stages_match_reference_modules(case):
    name, g, cfg, W, inputs, o32, _ = case
    s = int(g["stride"])
    for key, tap in [("ref32_projected", "projected"), ("ref32_llm_in", "llm_in"), ("ref32_hidden_1", "hidden.1"),
                     ("ref32_hidden_12", "hidden.12"), ("ref32_hidden_24", "hidden.24")]:
        got = o32[tap].float().reshape(-1)[::s].numpy()
        assert got.shape == g[key].shape
        assert _rel(got, g[key]) <= 1e-5, (name, key, _rel(got, g[key]))


def test_policy_head_bit_exact(case):
    name, g, cfg, W, (pix, ids, prop), _, o16 = case
    head = O.policy_head(o16["multi"], prop.to(torch.bfloat16), W, cfg, cfg.num_patches, torch.bfloat16)
    assert np.array_equal(head.float().numpy(), g["ref_head_on_oracle_multi"]), name


def test_end_to_end_vs_reference_bf16(case):
    name, g, cfg, W, inputs, o32, o16 = case
    st = STATS["synthetic"]["action"]
    ref = g["ref_actions"]
    for o, tol in ((o32, 4e-2), (o16, 5e-2)):
        un = O.unnormalize(o["normalized"].numpy().astype(np.float64), st["q99"], st["q01"], np.array(st["mask"]))
        assert un.shape == ref.shape
        assert np.abs(un - ref).max() <= tol, (name, np.abs(un - ref).max())
    ha = torch.from_numpy(g["ref_last_ha"]).view(torch.bfloat16).float().numpy()
    assert _rel(o32["last_ha"].float().numpy(), ha) <= 3e-2


def test_unnormalize_matches_reference_formula():
    """MP:799-803 with mask; masked-out dimension (gripper) passes through."""
    st = STATS["synthetic"]["action"]
    a = np.linspace(-1, 1, 56).reshape(8, 7)
    un = O.unnormalize(a, st["q99"], st["q01"], np.array(st["mask"]))
    hi, lo = np.array(st["q99"]), np.array(st["q01"])
    assert np.array_equal(un[:, 6], a[:, 6])
    assert np.allclose(un[:, :6], (0.5 * (a + 1) * (hi - lo + 1e-8) + lo)[:, :6], rtol=0, atol=0)


def test_reference_quirks_hold_in_oracle(case):
    """SURVEY 8a-7 / 8c invariants: h_t = [tok0, patch0..patch_{NP-2}], h_a = [last prompt token, AQ0..AQ62];
    causal mode: AQ63 and the stop token never influence the result; base variant rows are identical."""
    name, g, cfg, W, (pix, ids, prop), o32, _ = case
    NP, L = cfg.num_patches, ids.shape[1]
    multi = o32["multi"]
    assert multi.shape[1:] == (25, NP + 64, 896)
    assert torch.equal(multi[:, 3, 0], o32["hidden.3"][:, 0])
    assert torch.equal(multi[:, 3, NP], o32["hidden.3"][:, NP + L - 1])
    if not cfg.pro:
        n = o32["normalized"]
        assert all(torch.equal(n[:, 0], n[:, t]) for t in range(1, cfg.chunk_len))
    W2 = dict(W)
    aq = W2["vla.action_queries.weight"].clone()
    aq[63] += 1.0
    W2["vla.action_queries.weight"] = aq
    o2 = O.predict_action_batch(W2, cfg, pix, ids, prop, torch.float32)
    assert torch.equal(o2["normalized"], o32["normalized"])
