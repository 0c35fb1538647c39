"""World-size-2 gloo test (CPU) of the tensor-parallel host logic: the shard plan reproduces the full layer when
partial results are all-reduced, the vocab-split arg max merges with the first-maximum rule, and bad sizes are refused."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simplellminference_b200.config import PRESETS, ModelShape
from simplellminference_b200 import tp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _layer_reference(ms, W, x):
    """One transformer layer's matmul structure (no attention mixing: att := q, enough to exercise every matrix)."""
    q = W["wq"] @ x
    h = x + W["wo"] @ q
    s = (1.0 / (1.0 + np.exp(-(W["gate"] @ h)))) * (W["up"] @ h)
    y = W["down"] @ s + h
    return y, W["cls"] @ y


def _worker(rank, world, port, seed, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = PRESETS["tiny_gqa"]
    rng = np.random.default_rng(seed)
    d, I, V, kv = ms.hidden, ms.inter, ms.vocab, ms.kv_hidden
    W = {"wq": rng.standard_normal((d, d)), "wk": rng.standard_normal((kv, d)), "wv": rng.standard_normal((kv, d)),
         "wo": rng.standard_normal((d, d)), "up": rng.standard_normal((I, d)), "gate": rng.standard_normal((I, d)),
         "down": rng.standard_normal((d, I)), "cls": rng.standard_normal((V, d))}
    W = {k: v / np.sqrt(v.shape[1]) for k, v in W.items()}
    x = rng.standard_normal(d)
    plan = tp.shard_plan(ms, rank, world)
    loc = {k: W[k][plan[k].rows, plan[k].cols] for k in W}

    def allreduce(v):
        t = torch.from_numpy(np.ascontiguousarray(v))
        dist.all_reduce(t)
        return t.numpy()

    q_loc = loc["wq"] @ x                                   # column parallel: local heads
    h = x + allreduce(loc["wo"] @ q_loc)                    # row parallel + all-reduce, residual after the reduction
    s_loc = (1.0 / (1.0 + np.exp(-(loc["gate"] @ h)))) * (loc["up"] @ h)
    y = allreduce(loc["down"] @ s_loc) + h
    logits_loc = loc["cls"] @ y
    pair = torch.tensor([float(logits_loc.max()), float(plan["local"]["vocab_first"] + int(np.argmax(logits_loc)))], dtype=torch.float64)
    pairs = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(pairs, pair)
    token = tp.merge_argmax([(float(p[0]), int(p[1])) for p in pairs])
    y_ref, logits_ref = _layer_reference(ms, W, x)
    ok = np.allclose(y, y_ref, rtol=1e-10, atol=1e-10) and token == int(np.argmax(logits_ref))
    assert loc["wk"].shape == (kv // world, d) and loc["wo"].shape == (d, d // world) and loc["down"].shape == (d, I // world)
    out_q.put((rank, bool(ok), token))
    dist.destroy_process_group()


def test_shard_plan_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 123, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert len({tok for _, _, tok in res}) == 1      # every rank agrees on the token


def test_merge_argmax_first_maximum_and_divisibility():
    assert tp.merge_argmax([(1.0, 700), (1.0, 5), (0.5, 0)]) == 5
    assert tp.merge_argmax([(-1.0, 3)]) == 3
    with pytest.raises(ValueError):
        tp.shard_plan(ModelShape(512, 32, 96, 96, 100, 16, 1, 3, 3), 0, 2)
    p0, p1 = tp.shard_plan(PRESETS["llama2-7b"], 0, 8), tp.shard_plan(PRESETS["llama2-7b"], 7, 8)
    assert p0["local"] == dict(q=512, kv=512, inter=1376, vocab=4000, vocab_first=0, heads=4, kv_heads=4)
    assert p1["wo"].cols == slice(3584, 4096) and p1["cls"].rows == slice(28000, 32000)


# ---- the tensor-parallel prefill exchange (csrc/prefill_tp.cu), emulated: row ownership, rank-order sums, handshake ------------------
def _rmsnorm_bf16(x, w, eps):
    """(x * inv) * w rounded to bf16 (RNE), as the exchange kernel and pf_rmsnorm do; returned as the uint16 bit patterns."""
    x = x.astype(np.float32)
    inv = np.float32(1.0) / np.sqrt(np.float32((x * x).sum(dtype=np.float32) / np.float32(x.size)) + np.float32(eps))
    y = ((x * inv).astype(np.float32) * w.astype(np.float32)).astype(np.float32)
    u = y.view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


class _FakeLib:
    """Stands in for libsllm_b200 in Engine.init_p2p: records what is exported / imported."""

    def __init__(self, rank, with_prefill_block):
        self.rank, self.with_pf, self.imported = rank, with_prefill_block, {}

    def _export(self, tag, buf):
        for i in range(64):
            buf[i] = (self.rank * 16 + tag + i) % 251
        return 0

    def sllm_engine_p2p_export(self, h, buf):
        return self._export(1, buf)

    def sllm_engine_prefill_p2p_export(self, h, buf):
        return self._export(2, buf) if self.with_pf else -2   # SLLM_ENOTSUP: this engine has no exchange block

    def sllm_engine_p2p_import(self, h, raw):
        self.imported["decode"] = bytes(raw)
        return 0

    def sllm_engine_prefill_p2p_import(self, h, raw):
        self.imported["prefill"] = bytes(raw)
        return 0

    def sllm_last_error(self):
        return b""


def _exchange_worker(rank, world, port, out_q):
    import types
    from simplellminference_b200.engine import Engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    # 1. handshake: both blocks' 64-byte handles reach every rank in rank order; an engine without a prefill block skips the second round
    for with_pf in (True, False):
        fake = types.SimpleNamespace(lib=_FakeLib(rank, with_pf), h=None, tp_size=world)
        Engine.init_p2p(fake, dist)
        want = lambda tag: b"".join(bytes((r * 16 + tag + i) % 251 for i in range(64)) for r in range(world))
        ok &= fake.lib.imported["decode"] == want(1) and fake.prefill_p2p == with_pf
        ok &= (fake.lib.imported.get("prefill") == want(2)) if with_pf else ("prefill" not in fake.lib.imported)
    # 2. the exchange itself for a ragged row count: every rank owns ceil(T / world) rows, sums ALL ranks' partial rows in rank order, adds
    #    its residual rows, normalises, and every rank ends up with the same bf16 rows as all-reduce + RMSNorm would give
    T, d, eps = 7, 64, 1e-5
    rng = np.random.default_rng(5)
    parts = rng.standard_normal((world, T, d)).astype(np.float32)      # every rank's partial-sum matrix (same seed everywhere: "peer memory")
    x = rng.standard_normal((T, d)).astype(np.float32)                 # residual stream before the exchange
    w = (1.0 + 0.02 * rng.standard_normal(d)).astype(np.float32)
    lo, hi = tp.prefill_rows(T, world, rank)
    mine = np.zeros((T, d), np.uint16)
    x_after = x.copy()
    for row in range(lo, hi):
        acc = parts[0, row].copy()
        for q in range(1, world):
            acc = (acc + parts[q, row]).astype(np.float32)
        x_after[row] = (x[row] + acc).astype(np.float32)
        mine[row] = _rmsnorm_bf16(x_after[row], w, eps)
    t = torch.from_numpy(mine.astype(np.int32))
    dist.all_reduce(t)                                                  # rows are disjoint: the sum is the "stores into every rank's buffer"
    got = t.numpy().astype(np.uint16)
    total = parts.sum(0, dtype=np.float32) if world == 2 else None      # two ranks: rank-order sum == the plain sum
    ref = np.stack([_rmsnorm_bf16((x[r] + total[r]).astype(np.float32), w, eps) for r in range(T)])
    ok &= np.array_equal(got, ref)
    owners = [tp.prefill_rows(T, world, r) for r in range(world)]
    ok &= owners[0][0] == 0 and owners[-1][1] == T and all(owners[i][1] == owners[i + 1][0] for i in range(world - 1))
    out_q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_prefill_exchange_handshake_and_row_ownership_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_prefill_rows_partition():
    for T, size in ((512, 8), (300, 8), (7, 2), (3, 8), (1, 4), (1024, 4)):
        spans = [tp.prefill_rows(T, size, r) for r in range(size)]
        assert spans[0][0] == 0 and max(hi for _, hi in spans) == T
        assert all(0 <= lo <= hi <= T for lo, hi in spans)
        assert sum(hi - lo for lo, hi in spans) == T and all(spans[i][1] == spans[i + 1][0] or spans[i + 1] == (T, T) for i in range(size - 1))
