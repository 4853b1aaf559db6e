"""Micro-batching of concurrent `/act` requests.

The reference server (vla-scripts/deploy.py:78-107) answers one request at a time: `get_server_action` ->
`get_vla_action` -> `predict_action` with batch 1.  The engine's forward is batched, so concurrent requests can share
one forward: `ActionBatcher` queues prepared observations, groups those with the same prompt length and
un-normalisation key (the engine takes one prompt length per call; the reference has no padding path either), waits at
most `max_wait_ms` for the batch to fill, runs ONE `predict_action_batch` and hands every caller its own chunk.

Only the queueing / grouping logic lives here; image preparation and tokenisation stay with the reference's
`get_vla_action` preamble (experiments/robot/openvla_utils.py:737-806).  The batcher is engine-agnostic (anything with
`predict_action_batch(input_ids, attention_mask, pixel_values, proprio, unnorm_key)` works), which is also how the
CPU tests drive it.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch


@dataclass
class _Request:
    input_ids: torch.Tensor      # (1, L) int64
    pixel_values: torch.Tensor   # (1, 6n, 224, 224)
    proprio: np.ndarray          # (P,)
    unnorm_key: Optional[str]
    done: threading.Event = field(default_factory=threading.Event)
    result: Any = None
    error: Optional[BaseException] = None
    t_submit: float = field(default_factory=time.perf_counter)


class ActionBatcher:
    """Thread-safe request queue in front of one engine.

    submit() blocks the calling (request) thread until its action chunk is ready; a single worker thread owns the
    engine (the engine is not thread-safe, like the reference model).  Batches are formed per (prompt length,
    unnorm_key) group in arrival order, up to `max_batch`, after at most `max_wait_ms` of waiting for company.
    """

    def __init__(self, engine, max_batch: int = 8, max_wait_ms: float = 2.0):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.engine, self.max_batch, self.max_wait = engine, int(max_batch), float(max_wait_ms) * 1e-3
        self._q: List[_Request] = []
        self._cv = threading.Condition()
        self._stop = False
        self.batches: List[int] = []          # sizes of the batches run so far (for tests / monitoring)
        self._worker = threading.Thread(target=self._run, name="vla-action-batcher", daemon=True)
        self._worker.start()

    # ------------------------------------------------------------------ client side
    def submit(self, input_ids, pixel_values, proprio, unnorm_key: Optional[str] = None, timeout: Optional[float] = None):
        """One observation in, one (T, A) float64 chunk out - the shape `get_vla_action` returns per request."""
        ids = torch.as_tensor(input_ids)
        if ids.dim() == 1:
            ids = ids[None]
        pix = torch.as_tensor(pixel_values)
        if pix.dim() == 3:
            pix = pix[None]
        if ids.shape[0] != 1 or pix.shape[0] != 1:
            raise ValueError("submit() takes ONE observation; concurrency is what forms the batch")
        req = _Request(ids, pix, np.asarray(proprio, dtype=np.float32).reshape(-1), unnorm_key)
        with self._cv:
            if self._stop:
                raise RuntimeError("batcher is closed")
            self._q.append(req)
            self._cv.notify_all()
        if not req.done.wait(timeout):
            raise TimeoutError("no action within the timeout")
        if req.error is not None:
            raise req.error
        return req.result

    def close(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._worker.join(timeout=5.0)

    # ------------------------------------------------------------------ worker side
    @staticmethod
    def _key(r: _Request) -> Tuple[int, Optional[str], Tuple[int, ...]]:
        return (int(r.input_ids.shape[1]), r.unnorm_key, tuple(r.pixel_values.shape[1:]))

    def _take_batch(self) -> List[_Request]:
        """Called with the lock held and a non-empty queue: the oldest request's group, in arrival order."""
        key = self._key(self._q[0])
        batch = [r for r in self._q if self._key(r) == key][: self.max_batch]
        taken = set(map(id, batch))
        self._q = [r for r in self._q if id(r) not in taken]
        return batch

    def _run(self) -> None:
        while True:
            with self._cv:
                while not self._q and not self._stop:
                    self._cv.wait()
                if self._stop and not self._q:
                    return
                # wait (bounded) for the oldest request's group to fill up
                deadline = self._q[0].t_submit + self.max_wait
                while not self._stop:
                    key = self._key(self._q[0])
                    if sum(1 for r in self._q if self._key(r) == key) >= self.max_batch:
                        break
                    left = deadline - time.perf_counter()
                    if left <= 0:
                        break
                    self._cv.wait(left)
                batch = self._take_batch()
            self._serve(batch)

    def _serve(self, batch: List[_Request]) -> None:
        try:
            ids = torch.cat([r.input_ids for r in batch], 0)
            pix = torch.cat([r.pixel_values for r in batch], 0)
            prop = np.stack([r.proprio for r in batch], 0)
            out = self.engine.predict_action_batch(ids, None, pix, prop, unnorm_key=batch[0].unnorm_key)
            actions = out[0]
            self.batches.append(len(batch))
            for i, r in enumerate(batch):
                r.result = np.asarray(actions[i])
        except BaseException as ex:  # every waiting caller gets the error, like the reference's per-request "error"
            for r in batch:
                r.error = ex
        finally:
            for r in batch:
                r.done.set()
