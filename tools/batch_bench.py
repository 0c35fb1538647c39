#!/usr/bin/env python
"""tools/batch_bench.py — aggregate decode tokens/s of the batched multi-sequence path (sllm_batch_*) against the number
of sequences stepping together, next to its HBM roofline.

  python tools/batch_bench.py [--config llama2-7b] [--batches 1,2,4,8,16] [--context 512] [--steps 64] [--json]
                              [--variants plain,graph,graph+rows4,graph+rows4+ksplit] [--exp-batches 8,16]

Every sequence first decodes `context` tokens (untimed), then `steps` tokens are timed with CUDA events on the batch's
stream; weights are synthetic (bf16 by default), the cache pages bf16. Algorithmic bytes of a step = every weight once
+ per sequence its K/V rows (sllm_batch_step_bytes), so the roofline fraction says how close the shared weight pass
stays to the HBM rate as the FMA work per weight grows with the batch. Variants other than "plain" switch on the
experimental development knobs (sllm_tune 5 / 6 / 7). With --json every variant prints one JSON line as soon as it is
done (bench.py attaches them to its own line as "batch_decode"; a variant that crashes loses only itself and its successors)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


VARIANTS = {   # name -> sllm_tune keys switched on (5 graph replay, 6 four-row GEMV body, 7 K-split down projection: all experimental)
    "plain": (),
    "graph": (5,),
    "graph+rows4": (5, 6),
    "graph+rows4+ksplit": (5, 6, 7),
    "tc": ("tc",),              # sllm_batch_set_tensor_cores: the projections as tcgen05 GEMMs over the live rows (bf16 operands)
    "tc+graph": ("tc", 5),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="llama2-7b")
    ap.add_argument("--wdtype", default="bf16", choices=["f32", "bf16", "int8"])
    ap.add_argument("--kvdtype", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--batches", default="1,2,4,8,16", help="sequence counts for the plain variant")
    ap.add_argument("--exp-batches", default="8,16", help="sequence counts for the experimental variants")
    ap.add_argument("--tc-batches", default="8,16,32,64", help="sequence counts for the tensor-core variants")
    ap.add_argument("--pf-bn", type=int, default=0, help="force the N tile of the tensor-core GEMMs (sllm_tune key 1; 0 = planner)")
    ap.add_argument("--pf-ksplit", type=int, default=-1, help="force the K split of the residual GEMMs (sllm_tune key 4; -1 = planner)")
    ap.add_argument("--variants", default="plain", help="comma list out of: " + ", ".join(VARIANTS))
    ap.add_argument("--context", type=int, default=512)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--page-len", type=int, default=64)
    ap.add_argument("--json", action="store_true", help="one JSON line per variant on stdout, printed as soon as the variant is done")
    args = ap.parse_args()

    import dataclasses
    import numpy as np
    import torch
    from simplellminference_b200 import _lib
    from simplellminference_b200.batch import BatchDecoder
    from simplellminference_b200.config import PRESETS, F32, BF16, INT8
    from simplellminference_b200.engine import Engine

    torch.cuda.set_device(0)
    wd = {"f32": F32, "bf16": BF16, "int8": INT8}[args.wdtype]
    kvd = {"f32": F32, "bf16": BF16}[args.kvdtype]
    variants = [v.strip() for v in args.variants.split(",") if v.strip()]
    for v in variants:
        if v not in VARIANTS:
            raise SystemExit(f"unknown variant {v!r}")
    need = args.context + args.steps + 8
    # the engine only lends its weights and RoPE tables here: keep its own (unused) dense cache small
    ms = dataclasses.replace(PRESETS[args.config], max_len=max(need, 64))
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = Engine(ms, w_dtype=wd, kv_dtype=kvd, stream=stream).load_synthetic(1234)
    lib = _lib.load()
    if args.pf_bn:
        _lib.check(lib.sllm_tune(1, args.pf_bn))
    if args.pf_ksplit >= 0:
        _lib.check(lib.sllm_tune(4, args.pf_ksplit))
    def run_variant(v):
        for key in (5, 6, 7):
            _lib.check(lib.sllm_tune(key, 1 if key in VARIANTS[v] else 0))
        rng = np.random.default_rng(1)
        rows = []
        for B in [int(b) for b in (args.batches if v == "plain" else args.tc_batches if "tc" in VARIANTS[v] else args.exp_batches).split(",")]:
            bd = BatchDecoder(eng, max_seqs=B, page_len=args.page_len, kv_dtype=kvd)
            if "tc" in VARIANTS[v]:
                bd.set_tensor_cores(True)
            for _ in range(B):
                bd.add([int(rng.integers(1, ms.vocab))])
            bd.step(args.context)                      # untimed: fills every sequence's pages up to the context
            torch.cuda.synchronize()
            bytes0 = bd.step_bytes()
            launches0 = bd.total_launches
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            bd.step(args.steps)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms_total = ev0.elapsed_time(ev1)
            step_bytes = 0.5 * (bytes0 + bd.step_bytes())   # mean over the timed positions (linear in the position)
            t_step = ms_total * 1e-3 / args.steps
            ach = step_bytes / t_step / 1e9
            rows.append({"sequences": B, "tokens_per_sec": B * args.steps / (ms_total * 1e-3), "ms_per_step": 1e3 * t_step,
                         "bytes_per_step": step_bytes, "achieved_gbs": ach, "frac_of_hbm_peak": ach / peak,
                         "kernels_per_step": (bd.total_launches - launches0) / args.steps,
                         "checksum": int(sum(int(bd.tokens(s)[-1]) for s in range(B)) % 1000003)})
            if not args.json:
                r = rows[-1]
                print(f"{v:20s} B={B:3d}  {r['tokens_per_sec']:9.1f} tok/s  {r['ms_per_step']:7.3f} ms/step  {r['achieved_gbs']:7.0f} GB/s "
                      f"({100 * r['frac_of_hbm_peak']:.1f} % of {peak:.0f})  {r['kernels_per_step']:.0f} kernels/step", flush=True)
            bd.close()
        if args.json:
            print(json.dumps({"variant": v, "tune_keys_on": list(VARIANTS[v]), "pf_bn": args.pf_bn, "pf_ksplit": args.pf_ksplit,
                              "what": f"sllm_batch_step: {args.config}-shaped, {args.wdtype} weights, {args.kvdtype} cache pages of {args.page_len}, every sequence at "
                                      f"positions {args.context}..{args.context + args.steps - 1}; aggregate tokens/s over the sequences; algorithmic bytes = weights "
                                      "once per step + each sequence's K/V rows",
                              "peak_gbs": peak, "peak_source": peak_src, "steps": args.steps, "by_batch": rows}), flush=True)

    failed = 0
    for v in variants:
        try:
            run_variant(v)
        except Exception as ex:   # an error that left the CUDA context usable must not cost the variants after it
            failed += 1
            if args.json:
                print(json.dumps({"variant": v, "error": repr(ex)[:300]}), flush=True)
            else:
                print(f"{v}: FAILED {ex!r}", flush=True)
    for key in (5, 6, 7):
        lib.sllm_tune(key, 0)
    eng.close()
    if failed:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
