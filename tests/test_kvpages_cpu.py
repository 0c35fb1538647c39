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
