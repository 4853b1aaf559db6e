"""ActionBatcher: grouping, ordering, bounded waiting and error fan-out, with a fake engine (CPU)."""
import threading
import time

import numpy as np
import pytest
import torch

from vla_adapter_b200.serving import ActionBatcher, TemporalEnsembler


class _Engine:
    def __init__(self, delay=0.0, fail_on=None):
        self.calls, self.delay, self.fail_on = [], delay, fail_on

    def predict_action_batch(self, input_ids, attention_mask, pixel_values, proprio, unnorm_key=None):
        self.calls.append((tuple(input_ids.shape), unnorm_key))
        if self.fail_on is not None and input_ids.shape[1] == self.fail_on:
            raise ValueError("boom")
        time.sleep(self.delay)
        # chunk value = first token id of the sample: lets each caller verify it got ITS result back
        a = np.stack([np.full((8, 7), float(input_ids[i, 0])) for i in range(input_ids.shape[0])])
        return a, a.astype(np.float32)


def _submit_many(b, specs):
    out, errs = {}, {}

    def go(i, L, key):
        try:
            out[i] = b.submit(torch.full((1, L), i, dtype=torch.int64), torch.zeros(1, 12, 224, 224), np.zeros(8), key)
        except Exception as e:  # noqa: BLE001
            errs[i] = e

    ts = [threading.Thread(target=go, args=(i, L, key)) for i, (L, key) in enumerate(specs)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(10)
    return out, errs


def test_concurrent_requests_share_a_forward_and_get_their_own_chunk():
    eng = _Engine(delay=0.05)
    b = ActionBatcher(eng, max_batch=4, max_wait_ms=200)
    out, errs = _submit_many(b, [(20, "libero")] * 4)
    b.close()
    assert not errs and len(out) == 4
    for i, a in out.items():
        assert a.shape == (8, 7) and a.dtype == np.float64 and np.all(a == i)
    assert b.batches == [4] and eng.calls == [((4, 20), "libero")]


def test_mixed_prompt_lengths_share_a_batch():
    """Chat prompts of 40-56 tokens (OU:783) batch together: right-padded ids + attention mask go to the engine."""
    seen = {}

    class _E(_Engine):
        def predict_action_batch(self, input_ids, attention_mask, pixel_values, proprio, unnorm_key=None):
            seen["mask"] = None if attention_mask is None else attention_mask.clone()
            seen["ids"] = input_ids.clone()
            return super().predict_action_batch(input_ids, attention_mask, pixel_values, proprio, unnorm_key)

    eng = _E(delay=0.02)
    b = ActionBatcher(eng, max_batch=3, max_wait_ms=300)
    out, errs = _submit_many(b, [(40, "k"), (56, "k"), (48, "k")])
    b.close()
    assert not errs and b.batches == [3] and eng.calls == [((3, 56), "k")]
    for i, a in out.items():
        assert np.all(a == i)
    lens = sorted(int(r.sum()) for r in seen["mask"])
    assert lens == [40, 48, 56]
    for row, m in zip(seen["ids"], seen["mask"]):
        n = int(m.sum())
        assert bool((m[:n] == 1).all()) and bool((m[n:] == 0).all()) and bool((row[n:] == 0).all())


def test_temporal_ensembler_reproduces_the_reference_schedule():
    """evaluate_calvin.py:408-489: three chunks predicted at steps 0, 1, 2 are averaged position-wise over 10 steps."""
    rng = np.random.default_rng(0)
    b0, b1, b2 = (rng.normal(size=(8, 7)) for _ in range(3))
    ens = TemporalEnsembler(8, 7, balancing_factor=None, max_chunks=3)
    got = []
    for step in range(10):
        if step < 3:
            ens.add((b0, b1, b2)[step])
        else:
            ens.step_without_prediction()
        got.append(ens.action())
    want = [b0[0], (b0[1] + b1[0]) / 2, (b0[2] + b1[1] + b2[0]) / 3]
    want += [(b0[t] + b1[t - 1] + b2[t - 2]) / 3 for t in range(3, 8)]       # the reference's t in range(2, 7) loop, +1
    want += [(b1[7] + b2[6]) / 2, b2[7]]
    # the reference loops t = 2..6 after its three explicit steps, i.e. chunk positions 3..7 of b0
    assert len(got) == len(want) == 10
    for g, w in zip(got, want):
        assert np.allclose(g, w)
    # the weighted form of vla_evaluation.py:205-217: newest chunk weighs exp(0), age i weighs exp(-0.1 i)
    ens = TemporalEnsembler(8, 7)
    ens.add(b0)
    ens.add(b1)
    w = np.exp(-0.1 * np.arange(2))
    assert np.allclose(ens.action(), (w[0] * b1[0] + w[1] * b0[1]) / w.sum())
    ens.reset()
    with pytest.raises(RuntimeError):
        ens.action()
    with pytest.raises(ValueError):
        ens.add(np.zeros((7, 7)))


def test_groups_by_prompt_length_and_key_and_caps_batch():
    eng = _Engine(delay=0.01)
    b = ActionBatcher(eng, max_batch=2, max_wait_ms=100, mix_lengths=False)
    specs = [(20, "a"), (20, "a"), (20, "a"), (31, "a"), (20, "b")]
    out, errs = _submit_many(b, specs)
    b.close()
    assert not errs and len(out) == 5
    for i, a in out.items():
        assert np.all(a == i)
    assert sorted(b.batches) == [1, 1, 1, 2]          # (20,a) x2 + (20,a) x1, (31,a), (20,b)
    assert all(n <= 2 for (n, _), _ in eng.calls)
    assert {(shape[1], key) for shape, key in eng.calls} == {(20, "a"), (31, "a"), (20, "b")}


def test_lone_request_is_not_held_longer_than_max_wait():
    b = ActionBatcher(_Engine(), max_batch=8, max_wait_ms=30)
    t0 = time.perf_counter()
    a = b.submit(torch.full((1, 12), 5, dtype=torch.int64), torch.zeros(1, 6, 224, 224), np.zeros(8))
    dt = time.perf_counter() - t0
    b.close()
    assert np.all(a == 5) and 0.02 <= dt < 0.5 and b.batches == [1]


def test_engine_error_reaches_every_caller_of_that_batch_only():
    eng = _Engine(fail_on=31)
    b = ActionBatcher(eng, max_batch=4, max_wait_ms=50, mix_lengths=False)
    out, errs = _submit_many(b, [(31, None), (31, None), (20, None)])
    b.close()
    assert set(errs) == {0, 1} and all(isinstance(e, ValueError) for e in errs.values())
    assert set(out) == {2} and np.all(out[2] == 2)
    with pytest.raises(RuntimeError):
        b.submit(torch.zeros(1, 4, dtype=torch.int64), torch.zeros(1, 6, 224, 224), np.zeros(8))
    with pytest.raises(ValueError):
        ActionBatcher(eng, max_batch=0)
