"""Host-side integer logic of predict_action: placeholder/stop ids, fake labels, the ActionQuery mask.

Mirrors (same names, same results, batched) the reference helpers
    OpenVLAForActionPrediction._prepare_input_for_action_prediction   modeling_prismatic.py:748-769
    OpenVLAForActionPrediction._prepare_labels_for_action_prediction  modeling_prismatic.py:771-784
    PrismaticForConditionalGeneration._process_action_masks           modeling_prismatic.py:456-461
    get_current_action_mask / get_next_actions_mask                   prismatic/training/train_utils.py:8-41
This is index handling, so results must be bit-exact (tests compare with torch.equal)."""
from __future__ import annotations

import torch

# prismatic/vla/constants.py:11-15
IGNORE_INDEX = -100
ACTION_TOKEN_BEGIN_IDX = 151386
STOP_INDEX = 2
NUM_TOKENS = 64


def prepare_input_for_action_prediction(input_ids: torch.Tensor, attention_mask: torch.Tensor):
    """Append NUM_TOKENS placeholder ids (value 1) and the stop id; extend the mask with ones."""
    B = input_ids.shape[0]
    placeholder = torch.ones((B, NUM_TOKENS), dtype=input_ids.dtype, device=input_ids.device)
    stop = torch.full((B, 1), STOP_INDEX, dtype=input_ids.dtype, device=input_ids.device)
    ext = torch.cat([input_ids, placeholder, stop], dim=-1)
    ones = torch.ones((B, ext.shape[-1] - attention_mask.shape[-1]), dtype=attention_mask.dtype,
                      device=attention_mask.device)
    return ext, torch.cat([attention_mask, ones], dim=-1)


def prepare_labels_for_action_prediction(labels: torch.Tensor, ext_ids: torch.Tensor) -> torch.Tensor:
    """Extend the all-IGNORE labels with ACTION_TOKEN_BEGIN_IDX+1 and put the stop id last."""
    B = labels.shape[0]
    extension = torch.full((B, ext_ids.shape[-1] - labels.shape[-1]), ACTION_TOKEN_BEGIN_IDX + 1,
                           dtype=labels.dtype, device=labels.device)
    labels = torch.cat([labels, extension], dim=-1)
    labels[:, -1] = STOP_INDEX
    return labels


def get_current_action_mask(token_ids: torch.Tensor, action_dim: int) -> torch.Tensor:
    cumsum = torch.cumsum(token_ids != IGNORE_INDEX, dim=1)
    mask = (1 <= cumsum) & (cumsum <= action_dim)
    return (token_ids > ACTION_TOKEN_BEGIN_IDX) & mask


def get_next_actions_mask(token_ids: torch.Tensor, action_dim: int) -> torch.Tensor:
    cumsum = torch.cumsum(token_ids != IGNORE_INDEX, dim=1)
    return (token_ids > ACTION_TOKEN_BEGIN_IDX) & (cumsum > action_dim)


def process_action_masks(labels: torch.Tensor, action_dim: int) -> torch.Tensor:
    return get_current_action_mask(labels, action_dim) | get_next_actions_mask(labels, action_dim)


def action_query_index(all_actions_mask: torch.Tensor) -> torch.Tensor:
    """Index form of _replace_input_embeddings (modeling_prismatic.py:442-452): column j of sample b gets
    ActionQuery row k if it is the k-th True column of the mask, else -1.  Raises like the reference's
    torch.stack would when rows have different True counts, or when the count is not NUM_TOKENS."""
    counts = all_actions_mask.sum(dim=1)
    if not bool((counts == NUM_TOKENS).all()):
        raise ValueError(f"every sample must have exactly {NUM_TOKENS} ActionQuery positions, got {counts.tolist()}")
    idx = torch.cumsum(all_actions_mask.to(torch.int32), dim=1) - 1
    return torch.where(all_actions_mask, idx, torch.full_like(idx, -1)).to(torch.int32)


def build(input_ids: torch.Tensor, attention_mask: torch.Tensor | None, action_dim: int):
    """predict_action's preamble (modeling_prismatic.py:922-937): returns ext_ids (B, L+65) int64,
    labels, all_actions_mask (bool) and aq_index (int32)."""
    if input_ids.dim() != 2:
        raise ValueError("input_ids must be (B, L)")
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    labels = input_ids.clone()
    labels[:] = IGNORE_INDEX
    ext, ext_mask = prepare_input_for_action_prediction(input_ids, attention_mask)
    labels = prepare_labels_for_action_prediction(labels, ext)
    mask = process_action_masks(labels, action_dim)
    return ext.to(torch.int64), labels, mask, action_query_index(mask), ext_mask
