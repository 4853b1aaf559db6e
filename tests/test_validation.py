"""Validation-time forward reuse: the host logic on CPU with a recording engine, and on the GPU against
the reference computation restated with the oracle (teacher-forced batch -> same chunk as predict_action)."""
import numpy as np
import pytest
import torch

from vla_adapter_b200 import tokens, validation as V


def _batch(lens, Lmax=None, seed=0, T=8, A=7):
    """A collator-style batch: input_ids = prompt | 64 action-token ids | stop, right-padded; labels IGNORE on the
    prompt and the padding."""
    g = torch.Generator().manual_seed(seed)
    B = len(lens)
    Lmax = Lmax or max(lens)
    W = Lmax + tokens.NUM_TOKENS + 1
    ids = torch.zeros(B, W, dtype=torch.int64)
    labels = torch.full((B, W), tokens.IGNORE_INDEX, dtype=torch.int64)
    for b, L in enumerate(lens):
        ids[b, :L] = torch.randint(3, 1000, (L,), generator=g)
        act = torch.randint(tokens.ACTION_TOKEN_BEGIN_IDX + 1, tokens.ACTION_TOKEN_BEGIN_IDX + 200, (tokens.NUM_TOKENS,), generator=g)
        ids[b, L:L + tokens.NUM_TOKENS] = act
        ids[b, L + tokens.NUM_TOKENS] = tokens.STOP_INDEX
        labels[b, L:L + tokens.NUM_TOKENS] = act
        labels[b, L + tokens.NUM_TOKENS] = tokens.STOP_INDEX
    return {"input_ids": ids, "labels": labels, "actions": torch.rand(B, T, A, generator=g) * 2 - 1}


class _Engine:
    def __init__(self):
        self.calls = []

    def predict_action_batch(self, input_ids, attention_mask, pixel_values, proprio, **kw):
        self.calls.append([r.clone() for r in input_ids])
        B = len(input_ids)
        n = np.stack([np.full((8, 7), 0.01 * float(r[0])) for r in input_ids]).astype(np.float32)
        return n.astype(np.float64), n


def test_prompt_lengths_and_layout_checks():
    b = _batch([9, 14, 11])
    assert V.prompt_lengths_from_labels(b["labels"]).tolist() == [9, 14, 11]
    bad = b["labels"].clone()
    bad[0, 9 + 10] = tokens.IGNORE_INDEX                      # a hole in the action positions
    with pytest.raises(ValueError):
        V.prompt_lengths_from_labels(bad)
    with pytest.raises(ValueError):
        V.prompt_lengths_from_labels(torch.full((2, 80), tokens.IGNORE_INDEX))


def test_forward_metrics_and_run_validation_mirror_the_reference_dict():
    eng = _Engine()
    b = _batch([6, 9], seed=1)
    b["pixel_values"], b["proprio"] = torch.zeros(2, 12, 224, 224), np.zeros((2, 8), np.float32)
    loss, m = V.forward_metrics(eng, b)
    assert [len(r) for r in eng.calls[0]] == [6, 9] and torch.equal(eng.calls[0][1], b["input_ids"][1, :9])
    pred = torch.stack([torch.full((8, 7), 0.01 * float(b["input_ids"][i, 0])) for i in range(2)]).to(torch.bfloat16).float()
    gt = b["actions"].to(torch.bfloat16).float()
    assert abs(loss - (pred - gt).abs().mean().item()) < 1e-7
    assert set(m) == {"loss_value", "curr_action_l1_loss", "next_actions_l1_loss"}
    assert abs(m["curr_action_l1_loss"] - (pred[:, 0] - gt[:, 0]).abs().mean().item()) < 1e-7
    avg = V.run_validation(eng, [b, b, b])
    assert avg["val_batches_count"] == 3 and abs(avg["loss"] - loss) < 1e-7 and "next_actions_l1_loss" in avg
    assert V.run_validation(eng, [b, b, b], val_time_limit=-1.0)["val_batches_count"] == 1
    with pytest.raises(ValueError):
        V.run_validation(eng, [])


@pytest.mark.gpu
def test_validation_forward_on_the_engine_matches_the_oracle():
    """finetune.py:398-409 gathers the action positions through the label mask; with the prompt lengths taken from the
    labels the engine reproduces that forward: L1 metrics against the oracle's chunk for each sample's own prompt."""
    from oracle import vla_oracle as O
    from vla_adapter_b200.engine import VLAEngine

    cfg = O.OracleConfig(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=1024, pro=False)
    W = O.make_weights(cfg, seed=19)
    lens = [12, 9, 15]
    b = _batch(lens, seed=3)
    pix, _, prop = O.make_inputs(cfg, len(lens), max(lens), seed=19)
    b["pixel_values"], b["proprio"] = pix, prop
    eng = VLAEngine(n_images=2, dino_depth=2, siglip_depth=2, vocab_size=1024, max_batch=3, max_prompt_len=max(lens))
    eng.load_flat(W)
    eng.finalize()
    loss, m = V.forward_metrics(eng, b)
    eng.close()
    want = []
    for i, L in enumerate(lens):
        o = O.predict_action_batch(W, cfg, pix[i:i + 1], b["input_ids"][i:i + 1, :L], prop[i:i + 1], torch.bfloat16)
        want.append(o["normalized"][0])
    pred = torch.stack(want).to(torch.bfloat16).float()
    gt = b["actions"].to(torch.bfloat16).float()
    ref_loss = (pred - gt).abs().mean().item()
    assert abs(loss - ref_loss) < 1e-2, (loss, ref_loss)
    assert abs(m["curr_action_l1_loss"] - (pred[:, 0] - gt[:, 0]).abs().mean().item()) < 1.5e-2
