"""Tensor-parallel shard plan (host-side restatement of csrc/engine.cu:shard_plan, used by tests and tooling).

Stored weights are [out][in] row-major (reference matmul_kernel.cpp:21). Column-parallel matrices (Wq, Wk, Wv, up,
gate, classifier) are split by OUTPUT rows; row-parallel ones (Wo, Wdown) by INPUT columns; partial results of the
row-parallel matmuls are summed over ranks (all-reduce). SURVEY.md §8e.
"""
from __future__ import annotations

from dataclasses import dataclass

from .config import ModelShape


@dataclass(frozen=True)
class Slice:
    rows: slice
    cols: slice


def check_divisible(shape: ModelShape, size: int) -> None:
    for name, v in (("heads", shape.heads), ("kv_heads", shape.kv_heads), ("inter", shape.inter), ("vocab", shape.vocab)):
        if v % size:
            raise ValueError(f"tensor parallel size {size} must divide {name}={v}")


def shard_plan(shape: ModelShape, rank: int, size: int) -> dict:
    """Per-layer slices of every matrix for `rank` of `size` (the same for every layer)."""
    check_divisible(shape, size)
    hd = shape.head_dim
    q = shape.heads // size * hd
    kv = shape.kv_heads // size * hd
    I = shape.inter // size
    V = shape.vocab // size
    full = slice(None)
    return {
        "wq": Slice(slice(rank * q, (rank + 1) * q), full),
        "wk": Slice(slice(rank * kv, (rank + 1) * kv), full),
        "wv": Slice(slice(rank * kv, (rank + 1) * kv), full),
        "wo": Slice(full, slice(rank * q, (rank + 1) * q)),
        "up": Slice(slice(rank * I, (rank + 1) * I), full),
        "gate": Slice(slice(rank * I, (rank + 1) * I), full),
        "down": Slice(full, slice(rank * I, (rank + 1) * I)),
        "cls": Slice(slice(rank * V, (rank + 1) * V), full),
        "local": dict(q=q, kv=kv, inter=I, vocab=V, vocab_first=rank * V, heads=shape.heads // size, kv_heads=shape.kv_heads // size),
    }


def merge_argmax(pairs):
    """(value, global index) per rank -> global arg max, first maximum wins (argmax.cpp:11)."""
    best_v, best_i = float("-inf"), 2**31 - 1
    for v, i in pairs:
        if v > best_v or (v == best_v and i < best_i):
            best_v, best_i = v, i
    return best_i


def prefill_rows(n_rows: int, size: int, rank: int) -> tuple[int, int]:
    """Prompt rows [lo, hi) that `rank` owns in the tensor-parallel prefill exchange (csrc/prefill_tp.cu): blocks of
    ceil(n_rows / size) rows in rank order; the last ranks may own fewer rows, or none."""
    per = (n_rows + size - 1) // size
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)
