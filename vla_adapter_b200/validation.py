"""Validation-time forward of the fine-tuning script on the engine (forward only).

`vla-scripts/finetune.py:run_forward_pass` (L1-regression branch, :404-447) runs the teacher-forced VLM forward on a
training batch, gathers the 25 hidden states at the vision rows and at the 64 action positions, calls the action head
and compares its chunk with the ground-truth actions (L1); `run_validation` (:605-685) averages those metrics over the
validation loader under `torch.no_grad()`.  On that path the action-token ids never matter - their embeddings are
overwritten by the ActionQuery table (modeling_prismatic.py:418-454) - and attention is causal, so the computation is
exactly `predict_action` on (prompt, images, proprio): what this module feeds to `VLAEngine.predict_action_batch`, with
the per-sample prompt lengths read from the batch's labels (the first non-IGNORE position).  Gradients (the training
half of run_forward_pass) are out of scope: the engine has no backward kernels.
"""
from __future__ import annotations

import time
from typing import Any, Dict, Iterable, Optional, Tuple

import numpy as np
import torch

from . import tokens


def prompt_lengths_from_labels(labels: torch.Tensor) -> torch.Tensor:
    """Per sample: number of prompt tokens = index of the first label that is not IGNORE_INDEX.  Checks the layout the
    reference's gather relies on (finetune.py:398-409 reshapes the masked states to (B, 1, NUM_TOKENS, D)): every
    sample has exactly NUM_TOKENS action positions, contiguous, right behind the prompt."""
    labels = torch.as_tensor(labels)
    if labels.dim() != 2:
        raise ValueError("labels must be (B, L)")
    live = labels != tokens.IGNORE_INDEX
    if not bool(live.any(dim=1).all()):
        raise ValueError("a sample has no action labels")
    first = live.int().argmax(dim=1)
    mask = tokens.process_action_masks(labels, action_dim=7)          # current | next: every action position
    counts = mask.sum(dim=1)
    if not bool((counts == tokens.NUM_TOKENS).all()):
        raise ValueError(f"every sample needs {tokens.NUM_TOKENS} action positions, got {counts.tolist()}")
    idx = torch.arange(labels.shape[1])[None]
    want = (idx >= first[:, None]) & (idx < (first + tokens.NUM_TOKENS)[:, None])
    if not bool((mask == want).all()):
        raise ValueError("action positions must be contiguous and follow the prompt directly")
    if bool((first < 1).any()):
        raise ValueError("empty prompt")
    return first.to(torch.int32)


def forward_metrics(engine, batch: Dict[str, Any]) -> Tuple[float, Dict[str, float]]:
    """The L1-regression branch of run_forward_pass (finetune.py:404-447) without autograd: returns (loss, metrics) with
    the reference's metric names.  `batch` has the collator's keys: input_ids, labels, pixel_values, proprio, actions
    (B, T, A) normalised ground truth.  The predicted chunk is rounded to bf16 like the reference's head output
    (action_heads.py:53, 81) and the ground truth is cast to bf16 (finetune.py:331); the means are taken in fp32."""
    ids = torch.as_tensor(batch["input_ids"]).to("cpu", torch.int64)
    lens = prompt_lengths_from_labels(batch["labels"])
    prompts = [ids[b, : int(lens[b])] for b in range(ids.shape[0])]
    pix = torch.as_tensor(batch["pixel_values"])
    prop = batch["proprio"]
    _, normalized = engine.predict_action_batch(prompts, None, pix, prop)[:2]
    pred = torch.from_numpy(np.asarray(normalized)).to(torch.bfloat16).float()
    gt = torch.as_tensor(batch["actions"]).to(torch.bfloat16).float()
    if pred.shape != gt.shape:
        raise ValueError(f"ground-truth actions {tuple(gt.shape)} do not match the engine's chunk {tuple(pred.shape)}")
    loss = (pred - gt).abs().mean().item()
    metrics = {"loss_value": loss,
               "curr_action_l1_loss": (pred[:, 0] - gt[:, 0]).abs().mean().item(),
               "next_actions_l1_loss": (pred[:, 1:] - gt[:, 1:]).abs().mean().item() if gt.shape[1] > 1 else 0.0}
    return loss, metrics


def run_validation(engine, val_dataloader: Iterable[Dict[str, Any]], val_time_limit: Optional[float] = None) -> Dict[str, float]:
    """finetune.py:605-685: average of the per-batch metrics (plus `loss` = `loss_value` and `val_batches_count`), cut
    short when `val_time_limit` seconds have passed.  Returns the dict the reference logs to W&B."""
    t0 = time.time()
    all_metrics = []
    for batch in val_dataloader:
        _, m = forward_metrics(engine, batch)
        m["loss"] = m["loss_value"]
        all_metrics.append(m)
        if val_time_limit is not None and time.time() - t0 > val_time_limit:
            break
    if not all_metrics:
        raise ValueError("empty validation loader")
    avg = {k: sum(m[k] for m in all_metrics if k in m) / len([m for m in all_metrics if k in m]) for k in all_metrics[0]}
    avg["val_batches_count"] = len(all_metrics)
    return avg
