"""GPU parity tests of the batched multi-sequence decode over a paged KV cache (sllm_batch_*).

The reference decodes one sequence at a time, so the oracle of a batch is the oracle run once per sequence: every
sequence of a batch — whatever its neighbours, its slot, its pages and the moment it was admitted — must produce the
token stream the CPU oracle produces for that sequence alone (IDENTICAL tokens), with final logits inside the decode
tolerance of tests/test_engine_gpu.py (3e-4 * max(1, max|logit|) for fp32 cache rows, 5e-3 with a bf16 cache).

(The file sorts last on purpose: it is the newest subsystem; a failure here must not hide the rest of the suite.)"""
import numpy as np
import pytest

from conftest import oracle_shape
from simplellminference_b200 import _lib
from simplellminference_b200.batch import BatchDecoder
from simplellminference_b200.config import F32, BF16, INT8, PRESETS, ModelShape
from simplellminference_b200.engine import Engine

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def logit_tol(want, kv_dtype):
    return (3e-4 if kv_dtype == F32 else 5e-3) * max(1.0, float(np.abs(want).max()))


def run_ragged(port, ms, wd, kvd, seed, prompts, joins, n_steps, page_len, max_seqs, n_pages=None, threads=1):
    """Sequence i is admitted before step joins[i] and then steps with the others until step n_steps. Returns the engine's
    batch so that callers can look further; asserts tokens and final logits against the oracle of each sequence alone."""
    shape = oracle_shape(ms)
    blob = port.fill_blob(shape, seed, wd, 64, threads=threads)
    eng = Engine(ms, w_dtype=wd, kv_dtype=kvd).load_synthetic(seed)
    bd = BatchDecoder(eng, max_seqs=max_seqs, page_len=page_len, n_pages=n_pages, kv_dtype=kvd)
    slots = {}
    for step in range(n_steps):
        for i, j in enumerate(joins):
            if j == step:
                slots[i] = bd.add(prompts[i])
        bd.step(1)
    for i, prompt in enumerate(prompts):
        n = n_steps - joins[i]                       # steps this sequence took = tokens after its prompt[0]
        want, want_l = port.model(shape, blob, threads=threads, kv_bf16=(kvd == BF16)).greedy(prompt, n + 1)
        got = bd.tokens(slots[i])
        assert bd.position(slots[i]) == n
        assert np.array_equal(got, want), (i, np.flatnonzero(got != want)[:5], got[:8], want[:8])
        err = float(np.abs(bd.logits(slots[i]) - want_l).max())
        assert err <= logit_tol(want_l, kvd), (i, err)
    return eng, bd, slots


@pytest.mark.parametrize("wd,kvd", [(F32, F32), (BF16, F32), (INT8, F32), (BF16, BF16)])
def test_ragged_batch_matches_oracle_per_sequence(port, wd, kvd):
    """Five sequences of different prompt lengths admitted at different steps (GQA shape, pages of 4 positions)."""
    ms = PRESETS["tiny_gqa"]
    prompts = [[1, 7, 300], [5], [9, 2, 44, 17, 3, 8], [100, 200], [3]]
    eng, bd, _ = run_ragged(port, ms, wd, kvd, 1234, prompts, joins=[0, 0, 3, 10, 25], n_steps=40, page_len=4, max_seqs=6)
    bd.close(); eng.close()


def test_head_dim_48_and_page_of_one_position(port):
    """MHA shape with head_dim 48 (not a power of two), fp32 everywhere, the smallest possible page."""
    ms = PRESETS["tiny_mha_hd48"]
    eng, bd, _ = run_ragged(port, ms, F32, F32, 42, [[5], [7, 8, 9], [250, 1]], joins=[0, 1, 2], n_steps=30, page_len=1, max_seqs=3)
    bd.close(); eng.close()


def test_more_sequences_than_one_launch_holds(port):
    """Eleven sequences: the GEMV launcher goes through the slots in groups of 8 + 3."""
    ms = PRESETS["tiny_gqa"]
    rng = np.random.default_rng(8)
    prompts = [rng.integers(1, ms.vocab, size=int(rng.integers(1, 6))).tolist() for _ in range(11)]
    eng, bd, _ = run_ragged(port, ms, BF16, F32, 7, prompts, joins=[0] * 11, n_steps=24, page_len=8, max_seqs=11)
    bd.close(); eng.close()


def test_retire_and_readmit_recycles_slots_and_pages(port):
    """A retired sequence's slot and pages go to the next admission; the survivors are undisturbed."""
    ms = PRESETS["tiny_gqa"]
    shape = oracle_shape(ms)
    blob = port.fill_blob(shape, 11)
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32).load_synthetic(11)
    # 3 slots, pages of 4, and only as many pages as the test needs at its fullest moment (3 sequences of <= 20 positions)
    bd = BatchDecoder(eng, max_seqs=3, page_len=4, n_pages=15, kv_dtype=F32)
    a, b, c = bd.add([3]), bd.add([4, 5]), bd.add([6, 7, 8])
    assert (a, b, c) == (0, 1, 2)
    bd.step(12)
    free_before = bd.free_pages
    want_b, _ = port.model(shape, blob).greedy([4, 5], 13)
    assert np.array_equal(bd.tokens(b), want_b)
    bd.remove(b)
    assert bd.position(b) == -1 and bd.free_pages == free_before + 3       # 12 positions = 3 pages came back
    with pytest.raises(_lib.SllmError):
        bd.tokens(b)
    d = bd.add([9, 9, 9, 9])
    assert d == b                                                          # lowest free slot
    bd.step(8)
    for slot, prompt, n in ((a, [3], 20), (c, [6, 7, 8], 20), (d, [9, 9, 9, 9], 8)):
        want, want_l = port.model(shape, blob).greedy(prompt, n + 1)
        assert np.array_equal(bd.tokens(slot), want), slot
        assert float(np.abs(bd.logits(slot) - want_l).max()) <= logit_tol(want_l, F32)
    # the pool is now short for 3 more pages: all or nothing, nothing enqueued, positions unchanged
    assert bd.free_pages == 15 - (5 + 5 + 2)
    bd.step(4)                                                             # a, c: 24 positions = 6 pages; d: 12 = 3 pages -> 15 used
    assert bd.free_pages == 0
    with pytest.raises(_lib.SllmError) as ei:
        bd.step(1)
    assert ei.value.code == _lib.ENOMEM and [bd.position(s) for s in (a, c, d)] == [24, 24, 12]
    bd.remove(a)
    bd.step(1)
    assert [bd.position(s) for s in (c, d)] == [25, 13]
    want, _ = port.model(shape, blob).greedy([6, 7, 8], 26)
    assert np.array_equal(bd.tokens(c), want)
    bd.close(); eng.close()


def test_batch_of_one_equals_the_engine(port):
    """A single sequence through the batched path against the per-kernel engine path. The batched GEMV accumulates in
    the order of the one-sequence kernel and, at position 0, attention is the identity on the one value row, so the
    first step agrees to rounding noise of the epilogues (1e-6 relative; in practice bit for bit); then the same tokens."""
    ms = PRESETS["tiny_gqa"]
    eng = Engine(ms, w_dtype=BF16, kv_dtype=F32).load_synthetic(5)
    bd = BatchDecoder(eng, max_seqs=2, page_len=16, kv_dtype=F32)
    s = bd.add([17])
    bd.step(1)
    got = bd.logits(s)
    want, nxt = eng.forward(17, 0)
    assert float(np.abs(got - want).max()) <= 1e-6 * max(1.0, float(np.abs(want).max()))
    assert int(bd.tokens(s)[0]) == nxt
    q_b, q_e = bd.buffer("query")[s].cpu().numpy(), eng.buffer("query").cpu().numpy()[:ms.hidden]
    assert float(np.abs(q_b - q_e).max()) <= 1e-6 * max(1.0, float(np.abs(q_e).max()))
    bd.step(30)
    assert np.array_equal(bd.tokens(s), eng.greedy([17], 32))
    bd.close(); eng.close()


def test_argument_and_state_errors(port):
    ms = PRESETS["tiny_gqa"]
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32).load_synthetic(1)
    mega = Engine(ms, w_dtype=F32, kv_dtype=F32, mega=True).load_synthetic(1)
    if mega.mode.startswith("megakernel"):
        with pytest.raises(_lib.SllmError) as ei:
            BatchDecoder(mega, max_seqs=2)
        assert ei.value.code == _lib.ENOTSUP
    mega.close()
    for bad in (dict(max_seqs=0), dict(max_seqs=65), dict(max_seqs=2, page_len=0, n_pages=4), dict(max_seqs=2, n_pages=0)):
        with pytest.raises(_lib.SllmError):
            BatchDecoder(eng, **bad)
    bd = BatchDecoder(eng, max_seqs=2, page_len=8, kv_dtype=F32)
    bd.step(3)                                   # no live sequence: nothing to do
    with pytest.raises(_lib.SllmError):
        bd.add([ms.vocab])                       # token outside the vocabulary
    with pytest.raises(_lib.SllmError):
        bd.add(list(range(1, ms.max_len + 2)))   # prompt longer than max_len
    s0, s1 = bd.add([1]), bd.add([2])
    with pytest.raises(_lib.SllmError) as ei:
        bd.add([3])
    assert ei.value.code == _lib.ESTATE
    with pytest.raises(_lib.SllmError):
        bd.logits(s0)                            # has not stepped yet
    bd.step(ms.max_len)                          # to the last position
    assert bd.position(s0) == ms.max_len
    with pytest.raises(_lib.SllmError) as ei:
        bd.step(1)
    assert ei.value.code == _lib.EINVAL
    with pytest.raises(_lib.SllmError):
        bd.remove(5)
    bd.remove(s0); bd.remove(s1)
    with pytest.raises(_lib.SllmError):
        bd.remove(s0)
    bd.close(); eng.close()


def test_step_bytes_share_the_weights(port):
    """B(p) of a batch = the one-sequence B(p) with the weight term counted once."""
    ms = PRESETS["tiny_gqa"]
    eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16).load_synthetic(1)
    bd = BatchDecoder(eng, max_seqs=4, page_len=8, kv_dtype=BF16)
    for p in ([1], [2], [3]):
        bd.add(p)
    bd.step(5)
    one = ms.step_bytes(5, BF16, BF16)            # the next step of every sequence is at position 5
    per_seq = 2 * ms.hidden + 2 * 2 * ms.layers * ms.kv_hidden * (5 + 1) + 2 * 2 * ms.layers * ms.kv_hidden   # embedding row, K/V rows read, row written
    weights = one - per_seq
    assert abs(bd.step_bytes() - (weights + 3 * per_seq)) <= 8
    bd.close(); eng.close()


# ---- the cases below first ran at the end of round 1 on the driver's B200 (all passed): plain tests since then, so a regression
# of the batched decoder turns the suite red.


def test_full_width_llama2_7b_two_layers_batch_of_eight(port):
    """The Llama-2-7B WIDTHS (d 4096, inter 11008, vocab 32000, 32 heads of 128) with 2 layers and eight sequences: the shapes
    where the staged activations need opt-in shared memory (8 x 4096 floats = 128 KB per CTA) and where the down projection
    (11008 inputs: 5 vectors fit) goes through the slots in groups of 5 + 3. fp32 cache pages, so that token identity does
    not hinge on bf16 rounding boundaries."""
    import dataclasses
    import os
    ms = dataclasses.replace(PRESETS["llama2-7b"], layers=2, max_len=64)
    rng = np.random.default_rng(4)
    prompts = [rng.integers(1, ms.vocab, size=int(rng.integers(1, 5))).tolist() for _ in range(8)]
    eng, bd, _ = run_ragged(port, ms, BF16, F32, 9, prompts, joins=[0, 0, 0, 1, 1, 2, 3, 5], n_steps=14, page_len=16, max_seqs=8,
                            threads=os.cpu_count() or 1)
    assert bd.step_bytes() > 0
    bd.close(); eng.close()


def run_teacher_forced(port, ms, wd, kvd, seed, firsts, joins, n_steps, checkpoints, page_len, max_seqs, threads=1):
    """Long runs: a greedy stream of hundreds of tokens meets a near-tie sooner or later (margins of 1e-3 on 300-token
    runs of the synthetic model), so here every sequence is FED the oracle's own stream (a prompt as long as the run) and
    the logits are compared at checkpoints: a wrong cache row, page, tile or split shows up in every later logit."""
    shape = oracle_shape(ms)
    blob = port.fill_blob(shape, seed, wd, 64, threads=threads)
    fed, want = [], []
    for i, first in enumerate(firsts):
        om = port.model(shape, blob, threads=threads, kv_bf16=(kvd == BF16))
        toks, at = [int(first)], {}
        for pos in range(n_steps - joins[i]):
            lg = om.forward(toks[pos], pos)
            if joins[i] + pos + 1 in checkpoints:
                at[joins[i] + pos + 1] = lg.copy()
            toks.append(int(np.argmax(lg)))
        fed.append(toks[:-1])
        want.append(at)
        om.close()
    eng = Engine(ms, w_dtype=wd, kv_dtype=kvd).load_synthetic(seed)
    bd = BatchDecoder(eng, max_seqs=max_seqs, page_len=page_len, kv_dtype=kvd)
    slots, step = {}, 0
    for cp in sorted(checkpoints):
        while step < cp:
            for i, j in enumerate(joins):
                if j == step:
                    slots[i] = bd.add(fed[i])
            nxt = min([j for j in joins if j > step] + [cp])
            bd.step(nxt - step)              # several steps per call: pages are taken ahead for all of them
            step = nxt
        for i in slots:
            w = want[i][cp]
            err = float(np.abs(bd.logits(slots[i]) - w).max())
            assert err <= logit_tol(w, kvd), (cp, i, err)
    for i in slots:
        got = bd.tokens(slots[i])
        assert np.array_equal(got[:-1], fed[i][1:]), i                  # fed verbatim
        assert int(got[-1]) == int(np.argmax(bd.logits(slots[i]))), i   # then the arg-max of its own logits
    return eng, bd


@pytest.mark.parametrize("heads,kv_heads", [(8, 8), (8, 2)])
def test_long_context_many_slots_split_kv(port, heads, kv_heads):
    """Contexts of up to 300 positions in 20 slots: the attention kernel runs with several KV splits per (slot, head) and
    their in-kernel merge (8 KV heads: 160 CTAs per split -> 2 splits of up to 150 positions = 3 tiles of 64 each, through
    pages of 16; 2 KV heads with 4 query heads each: 6 splits), and the GEMV launcher goes through the slots in groups of
    8 + 8 + 4. Sequences are admitted in two waves so positions differ by 40."""
    ms = ModelShape(1024, 64, 512, 64 * kv_heads, 1408, 330, 2, heads, kv_heads)
    rng = np.random.default_rng(heads * 100 + kv_heads)
    firsts = rng.integers(1, ms.vocab, size=20).tolist()
    joins = [0 if i % 2 == 0 else 40 for i in range(20)]
    eng, bd = run_teacher_forced(port, ms, BF16, F32, 31, firsts, joins, n_steps=300, checkpoints={70, 150, 230, 300}, page_len=16,
                                 max_seqs=20)   # oracle single-threaded: its matrices are too small for a thread per row block to pay
    bd.close(); eng.close()


@pytest.mark.parametrize("name", ["cfg1_stories15M", "cfg2_stories110M", "tiny_gqa", "tiny_gqa_bf16w", "tiny_gqa_int8w", "tiny_mha_hd48"])
def test_golden_streams_of_the_reference_inside_a_batch(golden_models, name):
    """The token streams and final logits recorded from the UNMODIFIED reference (tests/golden/models_ref.npz, the fixtures
    tests/test_engine_gpu.py holds every engine mode to) reproduced by a sequence that shares its steps with two others."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    prompt, n_total, wd = mg.MODEL_RUNS[name]
    ms = PRESETS[{"cfg1_stories15M": "stories15M", "cfg2_stories110M": "stories110M", "tiny_mha_hd48": "tiny_mha_hd48"}.get(name, "tiny_gqa")]
    eng = Engine(ms, w_dtype=wd, kv_dtype=F32, group=64).load_synthetic(mg.SEED)
    bd = BatchDecoder(eng, max_seqs=3, page_len=16, kv_dtype=F32)
    other = bd.add([7])
    gold = bd.add(prompt)
    bd.step(3)
    late = bd.add([9, 11])                    # joins while the golden sequence is under way
    bd.step(n_total - 1 - 3)
    want, want_l = golden_models[name + "/tokens"], golden_models[name + "/last_logits"]
    got = bd.tokens(gold)
    assert np.array_equal(got, want), (np.flatnonzero(got != want)[:5], got[:8], want[:8])
    err = float(np.abs(bd.logits(gold) - want_l).max())
    assert err <= logit_tol(want_l, F32), err
    assert bd.position(other) == n_total - 1 and bd.position(late) == n_total - 4
    bd.close(); eng.close()


def test_continuous_batching_matches_oracle_per_request(port):
    """scheduler.ContinuousBatcher over the real decoder: nine requests through three slots and a pool that cannot hold
    three full-length requests at once (admissions are deferred), chunks of 5 steps; every request = the oracle alone.
    Then the same with an EOS id: each result is the oracle's stream cut after the first generated EOS."""
    from simplellminference_b200.scheduler import ContinuousBatcher
    ms = PRESETS["tiny_gqa"]
    shape = oracle_shape(ms)
    blob = port.fill_blob(shape, 1234)
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32).load_synthetic(1234)
    bd = BatchDecoder(eng, max_seqs=3, page_len=4, n_pages=24, kv_dtype=F32)
    rng = np.random.default_rng(12)
    reqs = [(rng.integers(1, ms.vocab, size=int(rng.integers(1, 8))).tolist(), int(rng.integers(5, 36))) for _ in range(9)]
    want = [port.model(shape, blob).greedy(p, len(p) + m)[0] for p, m in reqs]
    cb = ContinuousBatcher(bd, chunk=5)
    ids = [cb.submit(p, m) for p, m in reqs]
    out = cb.run()
    for rid, w in zip(ids, want):
        assert np.array_equal(out[rid], w), rid
    assert bd.free_pages == 24 and cb.stats.max_live <= 3 and cb.stats.admissions_deferred > 0
    eos = int(want[0][-1])
    cb = ContinuousBatcher(bd, eos_id=eos, chunk=5)
    ids = [cb.submit(p, m) for p, m in reqs]
    out = cb.run()
    for rid, w, (p, m) in zip(ids, want, reqs):
        hit = np.flatnonzero(w[len(p) - 1:] == eos)
        assert np.array_equal(out[rid], w[:len(p) - 1 + int(hit[0]) + 1] if hit.size else w), rid
    assert out[ids[0]][-1] == eos and bd.free_pages == 24
    bd.close(); eng.close()


def test_cpp_mirror_predict_batch(tmp_path):
    """model::LlamaModel::predict_batch of the C++ host mirror (waves of prompts through sllm_batch_*) against the oracle per prompt."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "_build", "test_batch_mirror")
    if not os.path.exists(exe):   # prebuilt by build(); a rebuild on the GPU box would recompile every CUDA file if the copy lost the mtimes
        for sub in (("simplellminference_b200", "csrc"), ("simplellminference_b200", "host"), ("tests", "cpp")):
            subprocess.run(["make", "-s", "-C", os.path.join(root, *sub)], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout


def test_graph_replay_matches_direct_launches():
    """Opt-in development knob sllm_tune(5, 1): the step's launch sequence depends on the live-slot count alone (tokens,
    positions and block tables are device memory), so it is captured into one CUDA graph per count and replayed. Same
    kernels, same arguments: tokens and logits must be bit-identical to the direct launches."""
    lib = _lib.load()
    ms = PRESETS["tiny_gqa"]
    eng = Engine(ms, w_dtype=BF16, kv_dtype=F32).load_synthetic(3)

    def run(graph):
        _lib.check(lib.sllm_tune(5, 1 if graph else 0))
        try:
            bd = BatchDecoder(eng, max_seqs=4, page_len=4, kv_dtype=F32)
            a, b = bd.add([1, 7, 300]), bd.add([5])
            bd.step(10)                       # two live slots: direct, capture, 8 replays
            c = bd.add([9, 2])
            bd.step(20)                       # three: a second graph
            bd.remove(b)
            bd.step(5)                        # a hole in the middle: same count, same graph, the slot is skipped on the device
            out = [(bd.tokens(s).copy(), bd.logits(s).copy()) for s in (a, c)]
            launches = bd.total_launches
            bd.close()
        finally:
            lib.sllm_tune(5, 0)
        return out, launches

    direct, n_direct = run(False)
    replay, n_replay = run(True)
    assert n_direct == n_replay                       # the accounting counts kernel nodes either way
    for (t0, l0), (t1, l1) in zip(direct, replay):
        assert np.array_equal(t0, t1) and np.array_equal(l0, l1)
    eng.close()


@pytest.mark.parametrize("preset,wd", [("tiny_gqa", F32), ("tiny_gqa", BF16), ("tiny_gqa", INT8), ("tiny_mha_hd48", F32)])
def test_four_row_gemv_body_is_bit_identical(preset, wd):
    """Opt-in development knob sllm_tune(6, 1): the GEMV body that takes two units (four weight rows) per warp at a time,
    used when three or more sequences share a launch. A lane visits its chunks in the same order, so tokens and logits must
    be bit-identical to the two-row body's (hd48 shape: unit counts that leave some warps with a single unit)."""
    lib = _lib.load()
    ms = PRESETS[preset]
    eng = Engine(ms, w_dtype=wd, kv_dtype=F32, group=64).load_synthetic(6)

    def run(rows4):
        _lib.check(lib.sllm_tune(6, 1 if rows4 else 0))
        try:
            bd = BatchDecoder(eng, max_seqs=6, page_len=8, kv_dtype=F32)
            slots = [bd.add([3 + i, 40 + i]) for i in range(3)]
            bd.step(6)                                   # three vectors: the 4-vector instantiation
            slots += [bd.add([100 + i]) for i in range(3)]
            bd.step(24)                                  # six: the 8-vector instantiation
            out = [(bd.tokens(s).copy(), bd.logits(s).copy()) for s in slots]
            bd.close()
        finally:
            lib.sllm_tune(6, 0)
        return out

    two, four = run(False), run(True)
    for (t0, l0), (t1, l1) in zip(two, four):
        assert np.array_equal(t0, t1) and np.array_equal(l0, l1)
    eng.close()


def test_sampling_per_slot_matches_the_single_sequence_sampler(port):
    """sllm_batch_set_sampling: a slot that samples draws exactly what predict.sample_ids draws for that sequence alone (same
    logits, same (seed, position) key), and its neighbours — greedy or sampling with other parameters — are not disturbed."""
    from simplellminference_b200.predict import sample_ids
    from simplellminference_b200.scheduler import ContinuousBatcher
    ms = PRESETS["tiny_gqa"]
    shape = oracle_shape(ms)
    blob = port.fill_blob(shape, 1234)
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32).load_synthetic(1234)
    bd = BatchDecoder(eng, max_seqs=4, page_len=8, kv_dtype=F32)
    g = bd.add([1, 7, 300])
    a = bd.add([5, 6])
    bd.set_sampling(a, 0.8, top_k=20, seed=11)
    c = bd.add([9])
    bd.set_sampling(c, 1.3, top_p=0.9, seed=5)
    bd.step(36)
    got = {s: bd.tokens(s).copy() for s in (g, a, c)}
    want_g, _ = port.model(shape, blob).greedy([1, 7, 300], 37)
    assert np.array_equal(got[g], want_g)
    assert np.array_equal(got[a], sample_ids(eng, [5, 6], 36, temperature=0.8, top_k=20, seed=11))
    assert np.array_equal(got[c], sample_ids(eng, [9], 36, temperature=1.3, top_p=0.9, seed=5))
    greedy_c, _ = port.model(shape, blob).greedy([9], 37)
    assert not np.array_equal(got[c], greedy_c)              # it did draw
    bd.set_sampling(a, 0.0)                                  # back to arg-max: from here on the slot is greedy again
    for s in (g, a, c):
        bd.remove(s)
    with pytest.raises(_lib.SllmError):
        bd.set_sampling(a, 1.0)                              # not in use any more
    # through the scheduler: the same request queued with its sampling parameters
    cb = ContinuousBatcher(bd, chunk=6)
    rid = cb.submit([5, 6], 35, sampling=dict(temperature=0.8, top_k=20, seed=11))
    rid2 = cb.submit([1, 7, 300], 34)
    out = cb.run()
    assert np.array_equal(out[rid], got[a]) and np.array_equal(out[rid2], want_g)
    bd.close(); eng.close()


@pytest.mark.parametrize("wd", [F32, BF16, INT8])
def test_split_down_projection_matches_oracle(port, wd):
    """Opt-in development knobs sllm_tune(6, 1) + (7, 1): with an intermediate size whose rows do not fit shared memory eight
    times (8192 floats: 7 fit), eight live sequences take the down projection with K cut in two over grid.y (eight half rows
    fit) and x = h + (p0 + p1). A different summation split, so the check is the oracle per sequence, not bit-identity."""
    lib = _lib.load()
    ms = ModelShape(512, 32, 128, 64, 8192, 48, 2, 4, 2)
    prompts = [[3 + i, 90 + 2 * i] for i in range(9)]
    for key in (6, 7):
        _lib.check(lib.sllm_tune(key, 1))
    try:
        eng, bd, _ = run_ragged(port, ms, wd, F32, 23, prompts, joins=[0] * 8 + [5], n_steps=20, page_len=8, max_seqs=9)   # seed: oracle margins >= 7e-3
        bd.close(); eng.close()
    finally:
        for key in (6, 7):
            lib.sllm_tune(key, 0)



def _gain1_blob(port, ms, seed):
    """bf16-exact synthetic blob with every projection scaled by 0.25 (tests/test_prefill_gpu.py::_blob explains why bf16-operand paths
    are measured on the gain-1 model: on the gain-4 one the operand rounding is amplified chaotically layer by layer)."""
    sh = oracle_shape(ms)
    blob = port.fill_blob(sh, seed, BF16)
    for seg in range(2, 9):
        off, cnt = port.segment(sh, seg)[:2]
        blob[off:off + cnt] *= 0.25
    return blob


@pytest.mark.parametrize("kvd", [F32, BF16])
def test_tensor_core_step_matches_oracle_per_sequence(port, kvd):
    """sllm_batch_set_tensor_cores: the projections of a step as tcgen05 GEMMs over the live rows. Seven sequences of different lengths,
    admitted at different steps, each FED its whole token list (so no feedback can hide or amplify an error); after the last fed token
    the logits of every sequence against the oracle's for that sequence alone within the bf16-operand tolerance (5e-3 of max|logit| on
    the gain-1 model; the fp32 GEMV path of the same batch is held to 3e-4 and asserted here as the A/B partner), and the K row of layer 0
    is the same in both modes up to the rounding of its inputs."""
    ms = ModelShape(2048, 64, 512, 256, 1408, 96, 3, 8, 4)
    shape = oracle_shape(ms)
    blob = _gain1_blob(port, ms, 31)
    rng = np.random.default_rng(12)
    lens = [40, 17, 33, 5, 28, 40, 11]
    joins = [0, 0, 2, 9, 9, 20, 30]
    fed = [rng.integers(1, ms.vocab, size=n, dtype=np.int32) for n in lens]
    want = []
    for ids in fed:
        om = port.model(shape, blob, threads=4, kv_bf16=(kvd == BF16))
        for p in range(len(ids) - 1):
            om.step(int(ids[p]), p)
        want.append(om.forward(int(ids[-1]), len(ids) - 1))
        om.close()
    worst = {}
    for tc in (False, True):
        eng = Engine(ms, w_dtype=BF16, kv_dtype=kvd).load_blob(blob)
        bd = BatchDecoder(eng, max_seqs=8, page_len=8, kv_dtype=kvd)
        if tc:
            bd.set_tensor_cores(True)
        slots, done = {}, {}
        for step in range(max(j + n for j, n in zip(joins, lens))):
            for i, j in enumerate(joins):
                if j == step:
                    slots[i] = bd.add(fed[i])
            bd.step(1)
            for i, s in slots.items():
                if i not in done and bd.position(s) == lens[i]:     # the step that consumed the last fed token
                    done[i] = bd.logits(s).copy()
                    assert np.array_equal(bd.tokens(s)[:-1], fed[i][1:]), i
                    bd.remove(s)
            slots = {i: s for i, s in slots.items() if i not in done}
        assert len(done) == len(fed)
        errs = [float(np.abs(done[i] - want[i]).max()) / max(1.0, float(np.abs(want[i]).max())) for i in range(len(fed))]
        worst[tc] = max(errs)
        for i, e in enumerate(errs):
            assert e <= (5e-3 if tc else (3e-4 if kvd == F32 else 5e-3)), (tc, i, e)
            srt = np.sort(want[i])
            if srt[-1] - srt[-2] > 4 * e * max(1.0, float(np.abs(want[i]).max())):
                assert int(np.argmax(done[i])) == int(np.argmax(want[i])), (tc, i)
        bd.close(); eng.close()
    print(f"\nbatched decode, {len(fed)} ragged sequences, kv={'f32' if kvd == F32 else 'bf16'}: max|dlogit|/max|logit| GEMV path {worst[False]:.1e}, tensor-core path {worst[True]:.1e}")


def test_tensor_core_step_refuses_other_weight_types():
    eng = Engine(PRESETS["tiny_gqa"], w_dtype=F32, kv_dtype=F32).load_synthetic(3)
    bd = BatchDecoder(eng, max_seqs=2, page_len=8, kv_dtype=F32)
    with pytest.raises(_lib.SllmError) as ei:
        bd.set_tensor_cores(True)
    assert ei.value.code == _lib.ENOTSUP
    bd.close(); eng.close()
