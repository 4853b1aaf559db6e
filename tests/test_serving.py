"""ActionBatcher: grouping, ordering, bounded waiting and error fan-out, with a fake engine (CPU)."""
import threading
import time

import numpy as np
import pytest
import torch

from vla_adapter_b200.serving import ActionBatcher


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


def test_groups_by_prompt_length_and_key_and_caps_batch():
    eng = _Engine(delay=0.01)
    b = ActionBatcher(eng, max_batch=2, max_wait_ms=100)
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
    b = ActionBatcher(eng, max_batch=4, max_wait_ms=50)
    out, errs = _submit_many(b, [(31, None), (31, None), (20, None)])
    b.close()
    assert set(errs) == {0, 1} and all(isinstance(e, ValueError) for e in errs.values())
    assert set(out) == {2} and np.all(out[2] == 2)
    with pytest.raises(RuntimeError):
        b.submit(torch.zeros(1, 4, dtype=torch.int64), torch.zeros(1, 6, 224, 224), np.zeros(8))
    with pytest.raises(ValueError):
        ActionBatcher(eng, max_batch=0)
