#!/usr/bin/env python
"""tools/fuse_bench.py — the bench's decode measurement for the EXPERIMENTAL megakernel with the down projection fused into the gate_up
phase (SLLM_ENGINE_MEGA_FUSE_DOWN), as a stand-alone process: same model, weights, prompt, warm-up and timed positions as bench.py's
main arm, so that its tokens can be compared with the main arm's (`--expect-checksum`) and its tokens/s read beside it.

  python tools/fuse_bench.py [--config llama2-7b] [--prompt-len 512] [--steps 128] [--warmup 8] [--expect-checksum N] [--plain]

One JSON line on stdout. bench.py runs it in a child process after its own timed regions (key "experiments"): whatever happens here
cannot touch the headline numbers."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="llama2-7b")
    ap.add_argument("--wdtype", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--kvdtype", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--prompt-len", type=int, default=512)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--expect-checksum", type=int, default=None, help="token_checksum of bench.py's main arm for the same arguments")
    ap.add_argument("--plain", action="store_true", help="the verified megakernel instead (a self-check of this script against bench.py)")
    args = ap.parse_args()

    import numpy as np
    import torch
    from bench import prompt_ids, peaks
    from simplellminference_b200.config import PRESETS, F32, BF16
    from simplellminference_b200.engine import Engine

    torch.cuda.set_device(0)
    ms = PRESETS[args.config]
    wd = {"f32": F32, "bf16": BF16}[args.wdtype]
    kvd = {"f32": F32, "bf16": BF16}[args.kvdtype]
    K, W, P = args.steps, max(args.warmup, 3), args.prompt_len
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = Engine(ms, w_dtype=wd, kv_dtype=kvd, stream=stream, mega=True, mega_fuse_down=not args.plain).load_synthetic(1234)
    mode = eng.mode
    ids = prompt_ids(P, ms.vocab)
    toks = eng.greedy(ids, P + 1)                       # the prompt token by token, like the main arm
    echo_ok = bool(toks.size == P and np.array_equal(toks[:P - 1], ids[1:]))
    eng.enqueue_steps(W)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    eng.enqueue_steps(K)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    tokens = eng.read_tokens(K)
    checksum = int(np.sum(tokens.astype(np.int64)) % 1000003)
    pos_first = P + W
    step_bytes = float(np.mean([eng.step_bytes(p) for p in range(pos_first, pos_first + K)]))
    peak, peak_src = peaks()
    ach = step_bytes / (ms_total * 1e-3 / K) / 1e9
    print(json.dumps({
        "mode": mode, "in_effect": mode == ("megakernel" if args.plain else "megakernel(fused-down)"), "tokens_per_sec": K / (ms_total * 1e-3),
        "ms_per_step": ms_total / K, "steps": K, "warmup": W, "positions": [pos_first, pos_first + K - 1], "achieved_gbs": ach,
        "frac_of_hbm_peak": ach / peak, "peak_source": peak_src, "prompt_echo_ok": echo_ok, "token_checksum": checksum,
        "tokens_match_main_arm": (checksum == args.expect_checksum) if args.expect_checksum is not None else None,
        "what": "same workload as the main arm's `value` (resident decode, CUDA events on the launching stream); the down projection's summation "
                "order is not fixed in this kernel, so identical tokens are expected but not guaranteed at near-ties"}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
