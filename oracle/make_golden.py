#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (see oracle/ref_shim.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python oracle/make_golden.py            # writes tests/golden/{libero_base,libero_pro,single_image}.npz

Each case fixes a seed; weights and inputs are the oracle's deterministic `make_weights` / `make_inputs`
(checksums of both are stored so that an RNG drift is detected instead of silently comparing different
problems).  What is stored, per case:
  * integer path (bit-exact pins): ext_ids, attention mask, labels, all_actions_mask produced by the
    reference's own _prepare_input_for_action_prediction / _prepare_labels_for_action_prediction /
    _process_action_masks (MP:748-784, 456-461 -> TU:8-41);
  * the reference's end-to-end bs=1 results in its own precision (bf16): un-normalised actions float64
    (B, T, A) and the returned last-layer ActionQuery states (B, 64, 896) (MP:892-972, one call per sample);
  * fp32 stage pins from the reference's own sub-modules run in fp32 (the whole path cannot run in fp32:
    AH:53 / MP:855 hard-cast to bf16): projector output, LLM hidden states 1 / 12 / 24 on a strided
    subsample (every `STRIDE`-th element, to keep the fixtures small);
  * the policy head alone (bf16, the only precision it supports) on the oracle's bf16 `multi` tensor.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402
from oracle import vla_oracle as O  # noqa: E402

STRIDE = 97
STATS = {"synthetic": {"action": {"q01": [-0.9, -0.8, -1.0, -0.5, -0.25, -0.1, 0.0],
                                  "q99": [0.9, 0.7, 1.0, 0.6, 0.35, 0.2, 1.0],
                                  "mask": [True, True, True, True, True, True, False]}}}

CASES = {
    # name: (pro, n_images, B, L, seed, dino_depth, siglip_depth)
    "libero_base": (False, 2, 2, 20, 0, 3, 3),
    "libero_pro": (True, 2, 1, 33, 1, 3, 3),
    # NOT a reference-pinned deployment case: with one image the reference's own glue would still build the head with
    # num_task_tokens = 512 (openvla_utils.py:505, AH:28) and mis-split h_t / h_a; the shim builds it with 256, which
    # is what the engine and the oracle implement.  Kept as a pin of the n = 1 arithmetic only.
    "single_image": (False, 1, 1, 31, 2, 3, 3),
    # the FULL-DEPTH architecture of BASELINE.json (DINOv2 24 / SigLIP 27 blocks, 24 LLM layers, 24 policy blocks,
    # L = 48), bs = 1, through the unmodified reference; only the vocabulary is cut (gather only)
    "libero_full_pro": (True, 2, 1, 48, 3, 24, 27),
}


def case_config(name):
    pro, n_images, B, L, seed, dd, sd = CASES[name]
    cfg = O.OracleConfig(n_images=n_images, dino_depth=dd, siglip_depth=sd, vocab_size=2048, pro=pro)
    return cfg, B, L, seed


def checksum(t: torch.Tensor) -> int:
    return zlib.crc32(t.contiguous().view(torch.uint8).numpy().tobytes()) if t.dtype != torch.bfloat16 else \
        zlib.crc32(t.contiguous().view(torch.int16).numpy().tobytes())


def weights_checksum(W) -> int:
    c = 0
    for k in sorted(W):
        c = zlib.crc32(k.encode(), c)
        c = zlib.crc32(W[k].contiguous().view(torch.int16).numpy().tobytes(), c)
    return c


def sub(t: torch.Tensor) -> np.ndarray:
    return t.float().reshape(-1)[::STRIDE].contiguous().numpy()


@torch.no_grad()
def make_case(name: str) -> dict:
    cfg, B, L, seed = case_config(name)
    W = O.make_weights(cfg, seed=seed)
    pix, ids, prop = O.make_inputs(cfg, B, L, seed=seed)
    out = {"w_crc": np.int64(weights_checksum(W)), "pix_crc": np.int64(checksum(pix)),
           "ids": ids.numpy(), "proprio": prop.numpy(), "stride": np.int64(STRIDE)}

    # ---- the reference in its own precision (bf16), end to end, bs=1 per call
    ns, vla, head, pp = R.build_reference(cfg, W, torch.bfloat16, norm_stats=STATS)
    acts, hids = R.reference_predict_action(ns, vla, head, pp, pix, ids, prop, "synthetic")
    out["ref_actions"] = np.stack(acts).astype(np.float64)
    out["ref_last_ha"] = torch.cat([h.reshape(1, 64, 896) for h in hids]).view(torch.int16).numpy()

    # ---- integer path through the reference's own helpers
    labels = ids.clone()
    labels[:] = ns.K.IGNORE_INDEX
    ext, att = vla._prepare_input_for_action_prediction(ids, torch.ones_like(ids))
    labels = vla._prepare_labels_for_action_prediction(labels, ext)
    mask = vla._process_action_masks(labels)
    out.update(ref_ext_ids=ext.numpy(), ref_attention_mask=att.numpy(), ref_labels=labels.numpy(),
               ref_mask=mask.numpy())

    # ---- policy head alone, on the oracle's bf16 gathered states
    o16 = O.predict_action_batch(W, cfg, pix, ids, prop, torch.bfloat16, keep_taps=True)
    head_out = head.predict_action(o16["multi"], proprio=prop, proprio_projector=pp)
    out["ref_head_on_oracle_multi"] = head_out.float().numpy()
    del vla, head, pp

    # ---- fp32 stage pins from the reference's sub-modules
    ns, vla32, _, _ = R.build_reference(cfg, W, torch.float32, norm_stats=STATS)
    projected = vla32._process_vision_features(pix.float(), None, False)
    emb = vla32.get_input_embeddings()(ext)
    aq = vla32.action_queries.weight.view(1, 64, 896).repeat(B, 1, 1)
    emb = vla32._replace_input_embeddings(emb.clone(), mask, aq)
    mm, mm_mask = vla32._build_multimodal_attention(emb, projected, att)
    lm = vla32.language_model(input_ids=None, attention_mask=mm_mask, inputs_embeds=mm, output_hidden_states=True,
                              return_dict=True)
    out["ref32_projected"] = sub(projected)
    out["ref32_llm_in"] = sub(mm)
    for i in (1, 12, 24):
        out[f"ref32_hidden_{i}"] = sub(lm.hidden_states[i])
    return out


def main():
    if not R.available():
        raise SystemExit("the reference is not mounted; golden vectors can only be generated in the build container")
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in (sys.argv[1:] or CASES):
        data = make_case(name)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), ref_actions[0,0] = {data['ref_actions'][0, 0]}")


if __name__ == "__main__":
    main()
