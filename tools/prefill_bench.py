#!/usr/bin/env python
"""tools/prefill_bench.py — time sllm_engine_prefill (tcgen05 GEMMs + block attention) on a shape preset.

  python tools/prefill_bench.py [--config llama2-7b] [--tokens 512] [--reps 5] [--bn 0|128|256]
Prints one JSON line: ms per prefill, prompt tokens/s, algorithmic TFLOP/s and the fraction of the
measured dense-bf16 peak (MEASURED_PEAKS.json). Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel list."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from simplellminference_b200 import _lib  # noqa: E402
from simplellminference_b200.config import PRESETS, BF16  # noqa: E402
from simplellminference_b200.engine import Engine  # noqa: E402


def prefill_flops(ms, T, tp=1):
    d, kv, I, L, V = ms.hidden, ms.kv_hidden, ms.inter, ms.layers, ms.vocab
    gemm = 2.0 * T * L * (2 * d * d + 2 * kv * d + 3 * I * d) + 2.0 * V * d
    attn = 2.0 * L * d * T * T
    return (gemm + attn) / tp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="llama2-7b")
    ap.add_argument("--tokens", type=int, default=512)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--ksplit", type=int, default=-1)
    ap.add_argument("--pair", type=int, default=-1, help="1/0 force/forbid the two-SM (cta_group::2) GEMM, -1 heuristic")
    a = ap.parse_args()
    ms = PRESETS[a.config]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    if a.bn:
        _lib.load().sllm_tune(1, a.bn)
    _lib.load().sllm_tune(2, a.pair)
    _lib.load().sllm_tune(3, a.pdl)
    _lib.load().sllm_tune(4, a.ksplit)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, stream=stream, mega=True).load_synthetic(1234)
    rng = np.random.default_rng(20260101)
    ids = rng.integers(1, ms.vocab, size=a.tokens, dtype=np.int32)
    ids[0] = 1
    for _ in range(2):
        eng.prefill(ids)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
    ev[0].record(stream)
    for r in range(a.reps):
        eng.prefill(ids)
        ev[r + 1].record(stream)
    torch.cuda.synchronize()
    times = [ev[r].elapsed_time(ev[r + 1]) for r in range(a.reps)]
    msec = float(np.median(times))
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1354.0)))
    except Exception:
        peak = 1354.0
    tf = prefill_flops(ms, a.tokens) / (msec * 1e-3) / 1e12
    print(json.dumps({"config": a.config, "tokens": a.tokens, "ms": round(msec, 3), "all_ms": [round(t, 3) for t in times],
                      "prompt_tokens_per_s": round(a.tokens / (msec * 1e-3), 1), "tflops": round(tf, 1), "peak_tflops": peak,
                      "tensor_pipe_frac": round(tf / peak, 4), "bn": a.bn, "pair": a.pair, "next_token": int(eng.read_tokens(1)[0])}))


if __name__ == "__main__":
    main()
