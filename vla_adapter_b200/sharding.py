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
    # A rank must reach the collective whatever happens to its shard, or every other rank waits in it forever:
    # an empty shard (n < world) contributes zero rows instead of calling `predict` with B = 0 (which the engine
    # rejects), and an exception in `predict` is carried THROUGH the gather - every rank learns that some rank failed
    # and raises together, after the collective.
    local, failure = None, None
    try:
        if hi > lo:
            local = predict(input_ids[lo:hi], pixel_values[lo:hi], proprio[lo:hi])
    except Exception as ex:  # noqa: BLE001 - re-raised below, after the collective
        failure = ex
    if world == 1:
        if failure is not None:
            raise failure
        return gather_chunks(local, n, group)
    bad, shape, dtype, device = _agree_on_chunk_layout(local, failure is not None, group)
    if failure is not None:
        raise failure
    if bad:  # nobody enters the chunk gather: the step is abandoned on every rank
        raise RuntimeError(f"predict_sharded: rank(s) {bad} failed in their shard; rank {rank} aborts the step with them")
    if local is None:
        local = torch.zeros((hi - lo,) + shape, dtype=dtype, device=device)
    return gather_chunks(local, n, group)


def _agree_on_chunk_layout(local: Optional[torch.Tensor], failed: bool, group=None):
    """One small all-gather in front of the chunk gather: every rank announces (failed, has chunk, T, A); ranks without
    a chunk (empty shard) take the layout of a rank that has one.  Returns (failed ranks, (T, A), dtype, device)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    device = local.device if local is not None else torch.device(
        "cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([1 if failed else 0, 0 if local is None else 1,
                         0 if local is None else local.shape[1], 0 if local is None else local.shape[2]],
                        dtype=torch.int64, device=device)
    every = torch.empty(world * 4, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, mine, group=group)
    every = every.cpu().view(world, 4)
    bad = [r for r in range(world) if every[r, 0]]
    have = [r for r in range(world) if every[r, 1]]
    t, a = (int(every[have[0], 2]), int(every[have[0], 3])) if have else (0, 0)
    return bad, (t, a), (local.dtype if local is not None else torch.float32), device
