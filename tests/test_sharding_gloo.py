"""N>1 path on CPU: world_size-2 gloo processes shard a batch by sample and all-gather the action chunks
(the only collective of the path).  The per-sample forward is a deterministic stand-in (the sharding layer
never looks inside it); what is checked is the partition, ordering, ragged shards and the collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vla_adapter_b200 import sharding


def test_shard_range_partitions_exactly():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 64, 65, 256):
            spans = [sharding.shard_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert list(sharding.shard_counts(world, n)) == sizes
    with pytest.raises(ValueError):
        sharding.shard_range(2, 2, 8)


def _fake_predict(ids, pix, prop):
    # per-sample function of all three inputs -> (B_local, 8, 7)
    s = ids.float().sum(1, keepdim=True) * 1e-6 + pix.float().mean(dim=(1, 2, 3)).unsqueeze(1) + prop.sum(1, keepdim=True)
    return (s.unsqueeze(-1) * torch.arange(1, 57).view(1, 8, 7)).float()


def _worker_failing(rank, world, port, n, ret):
    """A rank whose `predict` raises must not leave the others blocked in the collective."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ids = torch.randint(3, 1000, (n, 12), generator=g)
        pix = torch.randn(n, 12, 8, 8, generator=g)
        prop = torch.randn(n, 8, generator=g)

        def predict(i, p, q):
            if rank == 1:
                raise ValueError("engine rejected the shard")
            return _fake_predict(i, p, q)

        try:
            sharding.predict_sharded(predict, ids, pix, prop)
            ret[rank] = "returned"
        except ValueError as ex:
            ret[rank] = "ValueError" if rank == 1 else f"unexpected {ex}"
        except RuntimeError as ex:
            ret[rank] = "RuntimeError" if rank == 0 and "rank(s) [1] failed" in str(ex) else f"unexpected {ex}"
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ids = torch.randint(3, 1000, (n, 12), generator=g)
        pix = torch.randn(n, 12, 8, 8, generator=g)
        prop = torch.randn(n, 8, generator=g)
        calls = []

        def predict(i, p, q):
            calls.append(i.shape[0])
            return _fake_predict(i, p, q)

        out = sharding.predict_sharded(predict, ids, pix, prop)
        ref = _fake_predict(ids, pix, prop)
        lo, hi = sharding.shard_range(rank, world, n)
        ok = torch.equal(out, ref) and calls == ([hi - lo] if hi > lo else [])
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == world
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _spawn(target, n, world=2):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=target, args=(r, world, port, n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, "a rank hung or crashed"
    return dict(ret)


def test_predict_sharded_failure_on_one_rank_reaches_every_rank():
    assert _spawn(_worker_failing, 6) == {0: "RuntimeError", 1: "ValueError"}


def test_predict_sharded_fewer_samples_than_ranks():
    """n < world: the trailing ranks own no sample; they must skip `predict` and still take part in the gather."""
    assert _spawn(_worker, 1, world=3) == {0: True, 1: True, 2: True}
    assert _spawn(_worker, 2, world=3) == {0: True, 1: True, 2: True}


@pytest.mark.parametrize("n", [8, 7, 1])
def test_predict_sharded_world2_gloo(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}
