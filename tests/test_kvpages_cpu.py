"""Host side of the paged KV cache (sllm_kvpages_*, include/sllm_b200.h): pure bookkeeping, runs without a GPU.
The batched decoder (sllm_batch_*) hands out pages with exactly these rules."""
import numpy as np
import pytest

from simplellminference_b200 import _lib
from simplellminference_b200.batch import KvPages


def test_reserve_grows_by_whole_pages_and_is_idempotent():
    kp = KvPages(n_pages=10, page_len=4, max_seqs=3, max_pages_per_seq=5)
    assert kp.free == 10
    assert kp.reserve(0, 0) == 0 and kp.held(0) == 0
    assert kp.reserve(0, 1) == 1          # positions [0, 1) -> one page
    assert kp.reserve(0, 4) == 0          # still inside the first page
    assert kp.reserve(0, 5) == 1          # position 4 opens the second page
    assert kp.reserve(0, 3) == 0          # never shrinks
    assert kp.held(0) == 2 and kp.free == 8
    t = kp.table()
    assert t.shape == (3, 5)
    assert t[0, :2].tolist() == [0, 1] and (t[0, 2:] == -1).all() and (t[1:] == -1).all()   # page 0 is handed out first
    kp.close()


def test_pages_are_disjoint_across_sequences_and_recycled():
    kp = KvPages(n_pages=6, page_len=2, max_seqs=3, max_pages_per_seq=4)
    kp.reserve(0, 4); kp.reserve(1, 3); kp.reserve(2, 2)
    t = kp.table()
    used = t[t >= 0]
    assert sorted(used.tolist()) == list(range(5)) and kp.free == 1
    kp.release(1)
    assert kp.held(1) == 0 and kp.free == 3 and (kp.table()[1] == -1).all()
    kp.reserve(2, 8)                      # takes the recycled pages and the last free one
    t = kp.table()
    used = t[t >= 0]
    assert len(set(used.tolist())) == used.size == 6 and kp.free == 0
    assert t[2, 0] == 4                   # what a sequence already holds never moves
    kp.close()


def test_out_of_pages_is_all_or_nothing():
    kp = KvPages(n_pages=3, page_len=8, max_seqs=2, max_pages_per_seq=4)
    kp.reserve(0, 9)
    before = kp.table()
    with pytest.raises(_lib.SllmError) as ei:
        kp.reserve(1, 17)                 # needs 3, only 1 free
    assert ei.value.code == _lib.ENOMEM and "out of KV pages" in str(ei.value)
    assert kp.free == 1 and kp.held(1) == 0 and np.array_equal(kp.table(), before)
    kp.reserve(1, 8)
    assert kp.free == 0
    kp.close()


def test_argument_errors():
    kp = KvPages(n_pages=8, page_len=4, max_seqs=2, max_pages_per_seq=2)
    with pytest.raises(_lib.SllmError) as ei:
        kp.reserve(0, 9)                  # 3 pages > max_pages_per_seq
    assert ei.value.code == _lib.EINVAL
    for bad_seq in (-1, 2):
        with pytest.raises(_lib.SllmError):
            kp.reserve(bad_seq, 1)
        with pytest.raises(_lib.SllmError):
            kp.release(bad_seq)
    assert kp.held(5) == 0 and kp.free == 8
    kp.close()
    with pytest.raises(_lib.SllmError):
        KvPages(0, 4, 1, 1)


def test_random_walk_keeps_the_invariants():
    """Admissions, growth and retirements in random order: pages never shared, never lost."""
    rng = np.random.default_rng(3)
    n_pages, page_len, max_seqs, max_pages = 40, 16, 6, 12
    kp = KvPages(n_pages, page_len, max_seqs, max_pages)
    length = [0] * max_seqs
    for _ in range(2000):
        s = int(rng.integers(max_seqs))
        if rng.random() < 0.15:
            kp.release(s)
            length[s] = 0
        else:
            want = min(length[s] + int(rng.integers(1, 40)), max_pages * page_len)
            need = -(-want // page_len) - kp.held(s)
            if need <= kp.free:
                assert kp.reserve(s, want) == max(need, 0)
                length[s] = max(length[s], want)
            else:
                with pytest.raises(_lib.SllmError):
                    kp.reserve(s, want)
        t = kp.table()
        used = t[t >= 0]
        assert used.size == len(set(used.tolist())) == n_pages - kp.free
        for q in range(max_seqs):
            assert kp.held(q) == -(-length[q] // page_len) == int((t[q] >= 0).sum())
            assert (t[q, :kp.held(q)] >= 0).all()      # a sequence's pages are a prefix of its row
    kp.close()


def test_arena_bytes_and_page_budget_for_180gb():
    """sllm_batch_arena_bytes: the decoder's device footprint as host arithmetic (the layout pass of sllm_batch_create without an
    arena), and the page pool a budget of HBM can carry."""
    from simplellminference_b200.batch import arena_bytes, pages_that_fit
    from simplellminference_b200.config import BF16, F32, PRESETS
    ms = PRESETS["llama2-7b"]
    per_page_bf16 = 2 * ms.layers * ms.kv_heads * 64 * ms.head_dim * 2            # K and V of 64 positions, all layers and heads: 32 MiB
    assert per_page_bf16 == 32 << 20
    a1, a2 = arena_bytes(ms, 16, 64, 100, BF16), arena_bytes(ms, 16, 64, 101, BF16)
    assert a2 - a1 == per_page_bf16 and a1 % (1 << 20) == 0
    assert arena_bytes(ms, 16, 64, 100, F32) - a1 == 100 * per_page_bf16          # fp32 pages are twice as large
    fixed = a1 - 100 * per_page_bf16                                              # per-slot buffers + workspace: logits dominate
    assert 16 * ms.vocab * 4 <= fixed < 64 << 20
    # 180 GB of HBM minus 13.5 GB of bf16 weights: 64 slots of 4096 positions would need 4096 pages = 128 GiB -> they all fit
    budget = 180 * 10**9 - 14 * 10**9
    assert pages_that_fit(ms, 64, 64, budget, BF16) == 64 * 64
    assert pages_that_fit(ms, 64, 64, budget, F32) == (budget - (arena_bytes(ms, 64, 64, 1, F32) - 2 * per_page_bf16)) // (2 * per_page_bf16)
    small = pages_that_fit(ms, 64, 64, 10 * 10**9, BF16)
    assert 64 <= small < 64 * 64 and arena_bytes(ms, 64, 64, small, BF16) <= 10 * 10**9 < arena_bytes(ms, 64, 64, small + 1, BF16)
    assert pages_that_fit(ms, 64, 64, 1 << 30, BF16) == 0                         # not even one page per slot
    assert _lib.load().sllm_batch_arena_bytes(None, 1, 1, 1, BF16) == -1
    with pytest.raises(_lib.SllmError):
        arena_bytes(ms, 65, 64, 10, BF16)
