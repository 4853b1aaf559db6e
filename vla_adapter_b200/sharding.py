"""Sample sharding of a batched predict_action call across the GPUs of one box.

The path has no cross-sample interaction, so a batch partitions by sample: rank r of G owns a contiguous
slice, runs the full forward on its own weight replica (no collective in the forward), and the only
exchange is ONE all-gather of the (B_local, T, A) fp32 action chunks (224 bytes per sample), NCCL over
NVLink on GPUs, gloo in the CPU tests.  Precedent in the reference: vla-scripts/evaluate_calvin.py:221-222
slices the work by process id and :913-914 gathers one scalar per rank.  Unlike that slice
(`num_sequences // num_procs`, which silently drops the remainder), every sample is assigned here: the first
`n % world` ranks take one extra sample."""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch


def shard_range(rank: int, world: int, n: int) -> Tuple[int, int]:
    """[lo, hi) of the samples owned by `rank` when `n` samples are split over `world` ranks."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request rank={rank} world={world} n={n}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(world: int, n: int) -> Sequence[int]:
    return [shard_range(r, world, n)[1] - shard_range(r, world, n)[0] for r in range(world)]


def gather_chunks(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gathers the per-rank (B_local, T, A) chunks into the (n_total, T, A) batch order of `shard_range`.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_total:
            raise ValueError("single-process gather needs the whole batch")
        return local
    world = dist.get_world_size(group)
    counts = shard_counts(world, n_total)
    rank = dist.get_rank(group)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} samples, its shard has {counts[rank]}")
    cap = max(counts)
    pad = local
    if local.shape[0] < cap:
        pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(c == cap for c in counts):
        return out
    return torch.cat([out[r * cap: r * cap + counts[r]] for r in range(world)], dim=0)


def predict_sharded(predict: Callable[..., torch.Tensor], input_ids: torch.Tensor, pixel_values: torch.Tensor,
                    proprio: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None,
                    group=None) -> torch.Tensor:
    """Every rank passes the same global batch; rank r runs `predict` (e.g. VLAEngine.predict_action_batch's
    device form) on its shard only and receives all (B, T, A) chunks."""
    import torch.distributed as dist

    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = input_ids.shape[0]
    if pixel_values.shape[0] != n or proprio.shape[0] != n:
        raise ValueError("Non-homogenous batch of (text, image) input -- forward() does not support mixed batches!")
    lo, hi = shard_range(rank, world, n)
    local = predict(input_ids[lo:hi], pixel_values[lo:hi], proprio[lo:hi])
    return gather_chunks(local, n, group)
