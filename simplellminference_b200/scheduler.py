"""Continuous batching on top of the batched decoder (sllm_batch_*): requests wait in a queue, are admitted into free
slots as soon as the page pool can carry them to their end, step together, and retire individually.

Additive to the reference, whose ``predict`` (source/model/model.cpp:142-187) serves one prompt per call. Every request
is decoded exactly as ``predict`` would decode it alone — prompt echo, then first-max arg-max feedback — so its result
does not depend on what else is in flight (that is the parity property ``tests/test_zz_batch_gpu.py`` checks); the
scheduler only decides WHEN a request runs. Optional EOS stop like ``predict.predict_ids`` (the reference never stops).
That contract is the default fp32-activation decoder's; a decoder switched to the tensor-core step (``BatchDecoder(tensor_cores=True)``)
rounds activations to bf16 as GEMM operands: same independence of the neighbours, results within the bf16-operand tolerance instead.

Admission rule (no preemption, so it must be deadlock-free): a request is admitted only if the free pages cover its whole
life (prompt + new tokens) on top of what the requests already in flight may still take. Steps are enqueued in chunks:
``chunk`` device-resident steps per host round trip (fewer when a request would finish earlier)."""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field

import numpy as np


@dataclass
class _Request:
    rid: int
    prompt: np.ndarray
    total: int                      # positions this request decodes in all: its steps
    sampling: dict | None = None    # BatchDecoder.set_sampling keywords (temperature, top_k, top_p, seed); None = arg-max
    slot: int = -1
    tokens: np.ndarray | None = None
    done: bool = False
    stopped_by_eos: bool = False


@dataclass
class SchedulerStats:
    steps: int = 0                  # decoder steps enqueued
    slot_steps: int = 0             # sum over steps of live requests (= tokens produced, prompt echo included)
    max_live: int = 0
    admissions_deferred: int = 0    # times the queue head had to wait for pages or a slot
    history: list = field(default_factory=list)   # (step index, live requests) per chunk


class ContinuousBatcher:
    """``decoder``: a ``BatchDecoder`` (or anything with its add / step / tokens / remove / position / free_pages /
    max_seqs / page_len / n_pages / max_len interface). Not thread-safe: one host thread drives a decoder, like the reference's model."""

    def __init__(self, decoder, eos_id: int | None = None, chunk: int = 8):
        if chunk < 1:
            raise ValueError("chunk must be >= 1")
        self.dec, self.eos_id, self.chunk = decoder, eos_id, chunk
        self.waiting: deque[_Request] = deque()
        self.live: dict[int, _Request] = {}          # slot -> request
        self.finished: dict[int, _Request] = {}
        self.stats = SchedulerStats()
        self._next_id = 0

    # ---- requests ----
    def submit(self, prompt_ids, max_new_tokens: int, sampling: dict | None = None) -> int:
        """Queue a request: the prompt, then ``max_new_tokens`` generated tokens (arg-max, or drawn with ``sampling`` =
        the keywords of ``BatchDecoder.set_sampling``). Returns its id."""
        prompt = np.ascontiguousarray(prompt_ids, dtype=np.int32).reshape(-1)
        if prompt.size < 1 or max_new_tokens < 1:
            raise ValueError("a request needs a non-empty prompt and at least one new token")
        total = int(prompt.size) + int(max_new_tokens) - 1   # steps: positions 0 .. total-1 (the last prompt token's step yields the first new one)
        if self._pages_for(total) > self._pool_pages():
            raise ValueError(f"request of {total} positions can never fit the page pool")
        max_len = getattr(self.dec, "max_len", None)
        if max_len is not None and total > max_len:
            # refused HERE: admitted, it would make sllm_batch_step fail in the middle of a chunk ("cannot take N more steps") and
            # strand every other live request with its slots and pages held
            raise ValueError(f"request of {total} positions exceeds the decoder's max_len {max_len}")
        r = _Request(self._next_id, prompt, total, sampling)
        self._next_id += 1
        self.waiting.append(r)
        return r.rid

    def _pages_for(self, positions: int) -> int:
        return -(-positions // self.dec.page_len)

    def _pool_pages(self) -> int:
        return self.dec.n_pages

    def _owed_pages(self) -> int:
        """Pages the requests in flight may still take before they end."""
        owed = 0
        for r in self.live.values():
            owed += self._pages_for(r.total) - self._pages_for(self.dec.position(r.slot))
        return owed

    # ---- the loop ----
    def _admit(self) -> None:
        while self.waiting and len(self.live) < self.dec.max_seqs:
            r = self.waiting[0]
            if self._pages_for(r.total) + self._owed_pages() > self.dec.free_pages:
                self.stats.admissions_deferred += 1
                return                                   # FIFO: the head waits, nobody overtakes it
            self.waiting.popleft()
            r.slot = self.dec.add(r.prompt)
            if r.sampling:
                self.dec.set_sampling(r.slot, **r.sampling)
            self.live[r.slot] = r
        if self.waiting and len(self.live) >= self.dec.max_seqs:
            self.stats.admissions_deferred += 1

    def _retire(self, r: _Request, tokens: np.ndarray, eos: bool) -> None:
        r.tokens, r.done, r.stopped_by_eos = tokens, True, eos
        self.dec.remove(r.slot)
        del self.live[r.slot]
        self.finished[r.rid] = r

    def run_chunk(self) -> bool:
        """Admit what fits, enqueue one chunk of steps, retire what finished. False when nothing is left to do."""
        self._admit()
        if not self.live:
            if self.waiting:   # cannot happen with submit()'s check and an empty decoder; guard against a foreign decoder state
                raise RuntimeError("requests are waiting but none can be admitted into an empty decoder")
            return False
        n = min([self.chunk] + [r.total - self.dec.position(r.slot) for r in self.live.values()])
        self.dec.step(n)
        self.stats.steps += n
        self.stats.slot_steps += n * len(self.live)
        self.stats.max_live = max(self.stats.max_live, len(self.live))
        self.stats.history.append((self.stats.steps, len(self.live)))
        for r in list(self.live.values()):
            pos = self.dec.position(r.slot)
            finished = pos >= r.total
            if self.eos_id is None and not finished:
                continue                                 # nothing to look at yet: no host read for this request
            toks = self.dec.tokens(r.slot)               # tokens that followed positions 0 .. pos-1
            if self.eos_id is not None:
                gen = toks[r.prompt.size - 1:]           # generated part (the echo of the prompt is not searched)
                hit = np.flatnonzero(gen == self.eos_id)
                if hit.size:
                    self._retire(r, toks[:r.prompt.size - 1 + int(hit[0]) + 1], True)
                    continue
            if finished:
                self._retire(r, toks[:r.total], False)
        return bool(self.live or self.waiting)

    def run(self) -> dict[int, np.ndarray]:
        """Run until every submitted request has finished. Returns {request id: tokens}, where tokens = what
        ``Engine.greedy(prompt, len(prompt) + max_new_tokens)`` returns for that prompt alone (cut after EOS if asked)."""
        while self.run_chunk():
            pass
        return {rid: r.tokens for rid, r in self.finished.items()}


def predict_many(decoder, prompts, max_new_tokens: int, eos_id: int | None = None, chunk: int = 8):
    """``predict_ids`` for a list of prompts through one continuous batch; results in the order of ``prompts``."""
    cb = ContinuousBatcher(decoder, eos_id=eos_id, chunk=chunk)
    ids = [cb.submit(p, max_new_tokens) for p in prompts]
    out = cb.run()
    return [out[i] for i in ids]
