"""Micro-batching of concurrent `/act` requests.

The reference server (vla-scripts/deploy.py:78-107) answers one request at a time: `get_server_action` ->
`get_vla_action` -> `predict_action` with batch 1.  The engine's forward is batched, so concurrent requests can share
one forward: `ActionBatcher` queues prepared observations, groups those with the same un-normalisation key and image
layout, waits at most `max_wait_ms` for the batch to fill, runs ONE `predict_action_batch` and hands every caller its own
chunk.  Prompts of different lengths share a batch (chat prompts span 40-56 tokens, openvla_utils.py:783): they are
right-padded with an attention mask and the engine takes the per-sample lengths (`vla_predict`'s `prompt_len`);
`mix_lengths=False` restores one prompt length per batch.

`TemporalEnsembler` is the host-side action-chunk aggregation the CALVIN evaluation wraps around predict_action
(vla-scripts/vla_evaluation.py:205-217 defines the buffers, vla-scripts/evaluate_calvin.py:408-489 the schedule).

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

    def __init__(self, engine, max_batch: int = 8, max_wait_ms: float = 2.0, mix_lengths: bool = True):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.engine, self.max_batch, self.max_wait = engine, int(max_batch), float(max_wait_ms) * 1e-3
        self.mix_lengths = bool(mix_lengths)
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
    def _key(self, r: _Request) -> Tuple[int, Optional[str], Tuple[int, ...]]:
        return (0 if self.mix_lengths else int(r.input_ids.shape[1]), r.unnorm_key, tuple(r.pixel_values.shape[1:]))

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
            lens = [int(r.input_ids.shape[1]) for r in batch]
            L = max(lens)
            mask = None
            if min(lens) == L:
                ids = torch.cat([r.input_ids for r in batch], 0)
            else:  # right-pad: causal attention never lets a real token see the padding behind it
                ids = torch.zeros((len(batch), L), dtype=torch.int64)
                mask = torch.zeros((len(batch), L), dtype=torch.int64)
                for i, r in enumerate(batch):
                    ids[i, : lens[i]] = r.input_ids[0]
                    mask[i, : lens[i]] = 1
            pix = torch.cat([r.pixel_values for r in batch], 0)
            prop = np.stack([r.proprio for r in batch], 0)
            out = self.engine.predict_action_batch(ids, mask, pix, prop, unnorm_key=batch[0].unnorm_key)
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


class TemporalEnsembler:
    """Aggregation of overlapping action chunks, per environment (host side, a few hundred bytes).

    The CALVIN wrapper keeps, per episode, a (T, T, A) buffer of the last T chunks with a flipped upper-triangular
    validity mask and weights exp(-0.1 i) (vla-scripts/vla_evaluation.py:205-217; reset at :233-234): slot i holds the
    chunk predicted i steps ago, whose action for "now" is its i-th entry.  `add(chunk)` pushes the newest chunk,
    `action()` returns the weighted mean of every stored prediction for the current step.  With `weights=None` the mean
    is unweighted, which reproduces the hand-written schedule of vla-scripts/evaluate_calvin.py:408-489
    ((b0[t] + b1[t-1] + b2[t-2]) / 3 and its ramp-up / ramp-down) when at most `max_chunks` chunks are kept."""

    def __init__(self, chunk_len: int = 8, action_dim: int = 7, balancing_factor: Optional[float] = 0.1,
                 max_chunks: Optional[int] = None):
        self.T, self.A = int(chunk_len), int(action_dim)
        self.max_chunks = self.T if max_chunks is None else int(max_chunks)
        if not (1 <= self.max_chunks <= self.T):
            raise ValueError("max_chunks must be in [1, chunk_len]")
        self.weights = None if balancing_factor is None else np.exp(-balancing_factor * np.arange(self.T))
        self.reset()

    def reset(self) -> None:
        self.buffer = np.zeros((self.T, self.T, self.A))          # [age, step within chunk, action dim]
        self.valid = np.zeros((self.T,), dtype=bool)

    def add(self, chunk) -> None:
        chunk = np.asarray(chunk, dtype=np.float64)
        if chunk.shape != (self.T, self.A):
            raise ValueError(f"chunk must be ({self.T}, {self.A}), got {chunk.shape}")
        self.buffer[1:] = self.buffer[:-1].copy()
        self.valid[1:] = self.valid[:-1].copy()
        self.buffer[0], self.valid[0] = chunk, True
        self.valid[self.max_chunks:] = False

    def step_without_prediction(self) -> None:
        """One environment step with no new chunk (the tail of evaluate_calvin's schedule): everything ages by one."""
        self.buffer[1:] = self.buffer[:-1].copy()
        self.valid[1:] = self.valid[:-1].copy()
        self.valid[0] = False

    def action(self) -> np.ndarray:
        """Mean over the stored chunks of their prediction for the current step: chunk of age i contributes entry i."""
        ages = np.nonzero(self.valid)[0]
        if ages.size == 0:
            raise RuntimeError("no chunk stored")
        preds = self.buffer[ages, ages]                            # (n, A)
        w = np.ones(ages.size) if self.weights is None else self.weights[ages]
        return (preds * w[:, None]).sum(0) / w.sum()
