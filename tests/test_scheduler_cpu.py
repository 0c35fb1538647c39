"""Continuous batching (simplellminference_b200/scheduler.py) on the CPU: the scheduler is host logic over the
BatchDecoder interface, so here it drives a FAKE decoder — real page bookkeeping (sllm_kvpages_*, the same rules the
device path uses, including its all-or-nothing step), and a toy deterministic "model" whose next token depends only on
the sequence's own history. What is checked is scheduling: every request gets exactly what it would get alone, whatever
the chunk size, slot count or pool size; the pool is never overdrawn; admission is FIFO and deadlock-free."""
import numpy as np
import pytest

from simplellminference_b200 import _lib
from simplellminference_b200.batch import KvPages
from simplellminference_b200.scheduler import ContinuousBatcher, predict_many

VOCAB = 97


def toy_next(tok: int, pos: int) -> int:
    return (tok * 31 + pos * 7 + 3) % VOCAB


def alone(prompt, total):
    """What a request produces by itself: the tokens that follow positions 0..total-1 (prompt echo, then feedback)."""
    out, tok = [], int(prompt[0])
    for pos in range(total):
        nxt = int(prompt[pos + 1]) if pos + 1 < len(prompt) else toy_next(tok, pos)
        out.append(nxt)
        tok = nxt
    return np.array(out, np.int32)


class FakeDecoder:
    """The BatchDecoder interface over the real page bookkeeping and the toy model."""

    def __init__(self, max_seqs, page_len, n_pages, max_len=10_000):
        self.max_seqs, self.page_len, self.n_pages, self.max_len = max_seqs, page_len, n_pages, max_len
        self.kp = KvPages(n_pages, page_len, max_seqs, -(-max_len // page_len))
        self.seq = {}            # slot -> dict(prompt, pos, tok, hist)
        self.min_free = n_pages
        self.step_calls = 0

    def add(self, prompt):
        slot = min(s for s in range(self.max_seqs) if s not in self.seq)   # ValueError when full, like ESTATE
        self.seq[slot] = dict(prompt=[int(t) for t in prompt], pos=0, tok=int(prompt[0]), hist=[])
        return slot

    def remove(self, slot):
        del self.seq[slot]
        self.kp.release(slot)

    def position(self, slot):
        return self.seq[slot]["pos"] if slot in self.seq else -1

    @property
    def free_pages(self):
        return self.kp.free

    def step(self, n):
        self.step_calls += 1
        need = sum(max(0, -(-(q["pos"] + n) // self.page_len) - self.kp.held(s)) for s, q in self.seq.items())
        if need > self.kp.free:
            raise _lib.SllmError(_lib.ENOMEM, "out of KV pages")       # all or nothing, nothing enqueued
        for s, q in self.seq.items():
            self.kp.reserve(s, q["pos"] + n)
        self.min_free = min(self.min_free, self.kp.free)
        for _ in range(n):
            for q in self.seq.values():
                p = q["pos"]
                nxt = q["prompt"][p + 1] if p + 1 < len(q["prompt"]) else toy_next(q["tok"], p)
                q["hist"].append(nxt)
                q["tok"], q["pos"] = nxt, p + 1

    def tokens(self, slot):
        return np.array(self.seq[slot]["hist"], np.int32)


def make_requests(rng, n):
    return [(rng.integers(1, VOCAB, size=int(rng.integers(1, 12))).astype(np.int32), int(rng.integers(1, 40))) for _ in range(n)]


@pytest.mark.parametrize("max_seqs,page_len,n_pages,chunk", [(1, 4, 16, 8), (4, 4, 40, 1), (4, 8, 14, 5), (8, 16, 200, 64), (3, 1, 60, 7)])
def test_every_request_gets_what_it_would_get_alone(max_seqs, page_len, n_pages, chunk):
    rng = np.random.default_rng(max_seqs * 1000 + n_pages)
    reqs = make_requests(rng, 25)
    dec = FakeDecoder(max_seqs, page_len, n_pages)
    cb = ContinuousBatcher(dec, chunk=chunk)
    ids = [cb.submit(p, m) for p, m in reqs]
    out = cb.run()
    assert sorted(out) == ids and not cb.live and not cb.waiting and not dec.seq
    for rid, (p, m) in zip(ids, reqs):
        assert np.array_equal(out[rid], alone(p, len(p) + m - 1)), rid
        assert out[rid].size == len(p) + m - 1
    assert dec.free_pages == n_pages and dec.min_free >= 0          # everything returned, never overdrawn
    assert cb.stats.max_live <= max_seqs
    assert cb.stats.slot_steps >= sum(len(p) + m - 1 for p, m in reqs)   # chunks may run a finished-late request no further than its end
    assert cb.stats.slot_steps == sum(len(p) + m - 1 for p, m in reqs)


def test_small_pool_defers_but_never_deadlocks_and_keeps_fifo_order():
    dec = FakeDecoder(max_seqs=4, page_len=4, n_pages=12)       # each request below needs 5 pages: two fit, the third waits
    cb = ContinuousBatcher(dec, chunk=4)
    admitted = []
    real_add = dec.add
    dec.add = lambda prompt: (admitted.append(int(prompt[0])), real_add(prompt))[1]
    for first in (10, 20, 30, 40, 50):
        cb.submit([first], 20)
    out = cb.run()
    assert admitted == [10, 20, 30, 40, 50]                      # FIFO
    assert cb.stats.max_live == 2 and cb.stats.admissions_deferred > 0
    assert all(v.size == 20 for v in out.values()) and dec.free_pages == 12


def test_request_that_can_never_fit_is_refused_at_submit():
    cb = ContinuousBatcher(FakeDecoder(max_seqs=2, page_len=4, n_pages=3))
    with pytest.raises(ValueError, match="never fit"):
        cb.submit([1, 2, 3], 11)                                 # 13 positions = 4 pages > 3
    cb.submit([1, 2, 3], 10)                                     # 12 positions = 3 pages: exactly the pool
    with pytest.raises(ValueError):
        cb.submit([], 3)
    with pytest.raises(ValueError):
        cb.submit([1], 0)
    assert len(cb.run()) == 1


def test_eos_stops_a_request_and_frees_its_slot_early():
    prompt = np.array([5, 6, 7], np.int32)
    full = alone(prompt, 3 + 30 - 1)
    eos = int(full[10])                                          # a generated token (index >= len(prompt) - 1 = 2)
    first = 2 + int(np.flatnonzero(full[2:] == eos)[0])
    dec = FakeDecoder(max_seqs=2, page_len=4, n_pages=40)
    cb = ContinuousBatcher(dec, eos_id=eos, chunk=4)
    a = cb.submit(prompt, 30)
    b = cb.submit([eos, eos, 9], 6)                              # EOS inside a PROMPT is echo, not a stop
    out = cb.run()
    assert np.array_equal(out[a], full[:first + 1]) and out[a][-1] == eos and cb.finished[a].stopped_by_eos
    want_b = alone([eos, eos, 9], 3 + 6 - 1)
    cut = np.flatnonzero(want_b[2:] == eos)
    assert np.array_equal(out[b], want_b[:2 + int(cut[0]) + 1] if cut.size else want_b)
    assert dec.free_pages == 40


def test_requests_submitted_while_running_join_the_batch():
    dec = FakeDecoder(max_seqs=3, page_len=8, n_pages=30)
    cb = ContinuousBatcher(dec, chunk=3)
    a = cb.submit([1], 30)
    for _ in range(3):
        assert cb.run_chunk()
    b = cb.submit([2, 3], 10)                                    # arrives mid-flight: admitted at the next chunk boundary
    assert cb.run_chunk() and len(cb.live) == 2
    out = cb.run()
    assert np.array_equal(out[a], alone([1], 30)) and np.array_equal(out[b], alone([2, 3], 11))
    assert predict_many(FakeDecoder(2, 4, 20), [[1], [2, 3]], 5)[1].tolist() == alone([2, 3], 6).tolist()


def test_request_longer_than_max_len_is_refused_at_submit():
    """A request whose prompt + new tokens exceed the engine's max_len must be refused by submit(): admitted, it would fail inside
    a chunk of steps and leave every other live request un-retired."""
    dec = FakeDecoder(max_seqs=2, page_len=4, n_pages=100, max_len=32)
    cb = ContinuousBatcher(dec)
    ok = cb.submit([1, 2, 3], 10)
    with pytest.raises(ValueError, match="max_len"):
        cb.submit(list(range(1, 21)), 14)          # 20 + 14 - 1 = 33 positions > 32
    cb.submit(list(range(1, 21)), 13)              # 32 positions: exactly fits
    out = cb.run()
    assert ok in out and len(out) == 2
    assert dec.free_pages == dec.n_pages
