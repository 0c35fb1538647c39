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
