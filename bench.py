#!/usr/bin/env python
"""bench.py — decode tokens/s (batch 1, greedy) and fraction of the HBM roofline (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]                 our arm (CUDA, sm_100a)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] the reference's CPU path, host cores
  torchrun ... bench.py --gpus N ...                                   tensor parallel over N GPUs (one rank each)

Workload (config.workload): Llama-2-7B-shaped model (BASELINE.json configs[3], the configuration the metric
and the >=70 % target are quoted on; it fits one GPU), random-init synthetic weights (counter-based hash, bf16
storage), batch-1 greedy decode continuing a 512-token synthetic prompt. A "step" is ONE decoded token = one
pass of the whole forward hot path (all weights + the KV cache read once). Per-step working set (13.5 GB) is
far larger than the 126 MB L2, so no L2 flush is needed between steps (stated in config.l2).

One JSON line on stdout (rank 0). `value` = tokens/s with everything resident in HBM (CUDA-graph replays,
token feedback on the device); `e2e` = the same metric through the reference-facing call — forward(token,
pos) per token with host buffers: H2D of token+position, D2H of the next token, a host sync per token.
`roofline` = the dominant kernel (gate_up GEMV) timed alone with CUDA events, cycling over the layers so its
weights always come from HBM; `step_roofline` = whole-step algorithmic bytes B(p) / step time.
Secondary keys measured in child processes AFTER every timed region (N = 1; a failure there costs only its own key):
`batch_decode` (tools/batch_bench.py) and `experiments` (tools/fuse_bench.py, tools/microbench/fusion_probe); --no-batch skips them.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decode_tokens_per_sec_batch1_greedy"
UNIT = "tokens/s"
PROMPT_LEN = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="llama2-7b", help="shape preset (simplellminference_b200/config.py)")
    ap.add_argument("--wdtype", default="bf16", choices=["f32", "bf16", "int8"])
    ap.add_argument("--kvdtype", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--prompt-len", type=int, default=PROMPT_LEN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pdl", action="store_true")
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--no-mega", action="store_true", help="use the per-kernel fused CUDA-graph path instead of the persistent megakernel")
    ap.add_argument("--mega-ll", action="store_true", help="single GPU: the barrier-free {value,epoch}-word megakernel instead of the grid-barrier one")
    ap.add_argument("--mega-fuse-down", action="store_true", help="single GPU, experimental: the megakernel variant with the down projection fused into the gate_up phase (SLLM_ENGINE_MEGA_FUSE_DOWN)")
    ap.add_argument("--mega-v1", action="store_true", help="single GPU: the round-1 grid-barrier megakernel (five grid-wide dependency points per layer, deterministic sums) instead of "
                                                           "the default: csrc/megakernel2.cu with the fused down projection and the calibrated partition")
    ap.add_argument("--no-calibrate", action="store_true", help="skip sllm_engine_calibrate (per-CTA shares of every phase sized by the measured streaming rate of each SM)")
    ap.add_argument("--extras", action="store_true", help="after the timed regions, also run the secondary measurements in child processes (batched multi-sequence decode, "
                                                          "microbenchmarks, per-phase timelines); their JSON goes under batch_decode / experiments")
    ap.add_argument("--nccl", action="store_true", help="tensor parallel: NCCL all-reduce instead of the fused peer-memory one")
    ap.add_argument("--kernel-only", default=None, help="profiling aid: loop one fused kernel kind and exit")
    ap.add_argument("--no-batch", action="store_true", help="(kept for old command lines: the secondary measurements are opt-in now, see --extras)")
    return ap.parse_args()


def tensor_peak():
    """(dense bf16 TFLOP/s sustained, source) for the prefill tensor-pipe fraction."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1350.0, "fallback (B200_PROFILING.md sustained cuBLAS bf16)"


def prefill_flops(ms, T):
    """GEMM flops of T prompt rows + last-row classifier + causal attention."""
    d, kv, I, L, V = ms.hidden, ms.kv_hidden, ms.inter, ms.layers, ms.vocab
    return 2.0 * T * L * (2 * d * d + 2 * kv * d + 3 * I * d) + 2.0 * V * d + 2.0 * L * d * T * T


def peaks():
    """(HBM GB/s, source) — MEASURED_PEAKS.json when the driver wrote it, else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def prompt_ids(n, vocab):
    import numpy as np
    rng = np.random.default_rng(20260101)
    ids = rng.integers(1, vocab, size=n, dtype=np.int32)
    ids[0] = 1
    return ids


class ClockSampler:
    """nvidia-smi sampled while the GPU runs the measured kernel (B200_PROFILING.md 'clocks' line). The timed region of a default run is
    ~55 ms — shorter than nvidia-smi's start-up — so the sampler starts before the (untimed) prompt feed, which runs the very same
    kernel back to back for > 1 s, and runs through warm-up, timed steps and the end-to-end leg; every sample carries a timestamp and
    the record says how many fell inside the timed region itself."""

    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, f"/tmp/sllm_clocks_{os.getpid()}.csv"
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self, begin):
        import datetime
        if begin:
            self.t0 = datetime.datetime.now()
        else:
            self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, pw, inside = [], [], set(), [], 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f")
                if self.t0 and self.t1 and self.t0 <= ts <= self.t1:
                    inside += 1
            except ValueError:
                pass
            for nme, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_mhz_min": min(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": inside, "power_w_max": max(pw),
                "window": "prompt feed (same kernel, untimed) + warm-up + timed steps + end-to-end leg"}


# ------------------------------------------------------------------------------------------ reference arm --
def cpu_reference_sample(ms, pos, steps, warmup, want_fast=True, budget_s=240.0):
    """Times the reference's own CPU forward (oracle/_ref = the unmodified sources compiled by oracle/Makefile; falls back to the C
    port only if that library was never built) on REAL full-depth forwards of the full-shape model: `warmup` untimed + `steps` timed
    calls of LlamaModel::forward at position `pos` (bf16-rounded synthetic weights expanded to fp32: a 24.6 GiB blob for Llama-2-7B;
    the KV history is the zero-initialised cache — the arithmetic per token does not depend on the values). The reference is
    single-threaded by construction (no threads / OpenMP anywhere in it): cores = 1. A wall-clock budget cuts the run short on a slow
    host: the line then says how many steps were really timed."""
    import numpy as np
    from oracle import loader
    port = loader.Port()
    kind, flags, ref = "port", "gcc -O2 -ffp-contract=off (C restatement)", None
    if loader.have_ref():
        fast = want_fast and loader.cpu_supports_v3() and os.path.exists(loader.REF_FAST_SO)
        ref = loader.Ref(fast=fast)
        kind, flags = "reference", ref.flags
    cwd = os.getcwd()
    os.chdir("/tmp")  # LlamaModel::forward() opens layer_outputs_cpu.txt in the cwd on every call (model.cpp:42)
    t_start = time.perf_counter()
    try:
        shape = loader.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)
        blob = port.fill_blob(shape, 1234, loader.BF16)
        t_fill = time.perf_counter() - t_start
        m = ref.model(shape, blob) if ref else port.model(shape, blob)
        ts, tok, done_warm = [], 1, 0
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            m.step(tok, pos)
            dt = time.perf_counter() - t0
            if i >= warmup:
                ts.append(dt)
            else:
                done_warm += 1
            if time.perf_counter() - t_start > budget_s and ts:
                break
        m.close()
        del blob
    finally:
        os.chdir(cwd)
    t_token = statistics.median(ts)
    return {
        "value": 1.0 / t_token, "unit": UNIT, "cores": 1, "kind": kind,
        "sample": f"{len(ts)} timed + {done_warm} warm-up full-depth forwards ({ms.layers} layers, full widths, full classifier) at pos={pos}, median "
                  f"{t_token * 1e3:.0f} ms (min {min(ts) * 1e3:.0f}, max {max(ts) * 1e3:.0f}); weight blob {shape_gib(ms):.1f} GiB filled in {t_fill:.0f} s; "
                  f"build: {flags}; host has {os.cpu_count()} cores, the reference uses 1",
        "sec_per_token": t_token, "steps_timed": len(ts), "warmup_done": done_warm, "extrapolated": False,
    }


def shape_gib(ms):
    return 4.0 * ms.n_params() / 2**30


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from simplellminference_b200.config import PRESETS
    ms = PRESETS[args.config]
    base = cpu_reference_sample(ms, args.prompt_len, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": base["steps_timed"],
        "warmup": base["warmup_done"], "ms_per_step": 1e3 * base["sec_per_token"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": False,
        "config": workload_config(args, ms),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, ms):
    return {"workload": f"{args.config}-shaped (d={ms.hidden} L={ms.layers} H={ms.heads} KVH={ms.kv_heads} I={ms.inter} V={ms.vocab}), "
                        f"{args.wdtype} weights, {args.kvdtype} KV cache, batch-1 greedy decode continuing a {args.prompt_len}-token prompt",
            "prompt_len": args.prompt_len, "weights": args.wdtype, "kv_cache": args.kvdtype, "parallelism": f"tp{args.gpus}",
            "l2": "no flush: every step streams its whole working set (weights+KV >> 126 MB L2) from HBM"}


def batch_decode_sample(args):
    """Secondary figure: aggregate tokens/s of the batched multi-sequence decode (sllm_batch_*, tools/batch_bench.py) on the
    same model shape and context: plain (1 / 4 / 8 / 16 sequences), then the experimental launch / kernel variants (8 / 16).
    Runs in ONE child process after every timed region of this one and after its engine is gone; the child prints a JSON line
    per variant as it goes, so whatever happens there (error, time-out) costs only the variants not yet printed and never
    the headline numbers."""
    variants = ["plain", "tc+graph", "graph", "graph+rows4", "graph+rows4+ksplit"]   # tc+graph: the tensor-core step (8 .. 64 sequences)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "batch_bench.py"), "--config", args.config, "--wdtype", args.wdtype,
           "--kvdtype", args.kvdtype, "--context", str(args.prompt_len), "--batches", "1,4,8,16", "--exp-batches", "8,16",
           "--variants", ",".join(variants), "--steps", "64", "--json"]
    lines, note = _child_json_lines(cmd, 200)
    got = {d["variant"]: d for d in lines if isinstance(d, dict) and "variant" in d}
    out = got.get("plain") or {"error": note or "no output"}
    out["experimental"] = {v: got.get(v) or {"error": note or "not reached"} for v in variants[1:]}
    return out


SECONDARY_BUDGET_S = 240.0   # wall-clock budget shared by ALL secondary child processes: the worst case adds this much to the run, no more
_secondary_deadline = None


def _secondary_timeout(cap):
    """Seconds a secondary child may take: its own cap, cut to what is left of the shared budget (0 = skip it)."""
    global _secondary_deadline
    if _secondary_deadline is None:
        _secondary_deadline = time.monotonic() + SECONDARY_BUDGET_S
    left = _secondary_deadline - time.monotonic()
    return 0 if left < 20 else int(min(cap, left))


def _child_json_lines(cmd, timeout):
    """Runs a secondary measurement in a child process; returns (parsed JSON lines of its stdout, note on how it ended or None)."""
    stdout, note = "", None
    timeout = _secondary_timeout(timeout)
    if timeout == 0:
        return [], "skipped: the shared time budget of the secondary measurements is spent"
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        stdout = r.stdout or ""
        if r.returncode != 0:
            note = f"exit {r.returncode}: {(r.stderr or '').strip()[-300:]}"
    except subprocess.TimeoutExpired as ex:
        stdout = ex.stdout.decode(errors="replace") if isinstance(ex.stdout, bytes) else (ex.stdout or "")
        note = f"timed out after {timeout} s"
    except Exception as ex:
        note = repr(ex)[:300]
    out = []
    for ln in stdout.splitlines():
        try:
            out.append(json.loads(ln))
        except Exception:
            continue
    return out, note


def _child_text(cmd, timeout, max_lines=60):
    """Like _child_json_lines for a tool that prints a short human-readable report: its stdout lines (or the reason there are none)."""
    timeout = _secondary_timeout(timeout)
    if timeout == 0:
        return ["skipped: the shared time budget of the secondary measurements is spent"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        lines = [ln.rstrip()[:400] for ln in (r.stdout or "").splitlines() if ln.strip()][:max_lines]
        if r.returncode != 0:
            lines.append(f"exit {r.returncode}: {(r.stderr or '').strip()[-300:]}")
        return lines
    except subprocess.TimeoutExpired as ex:
        partial = ex.stdout.decode(errors="replace") if isinstance(ex.stdout, bytes) else (ex.stdout or "")
        return [ln.rstrip()[:400] for ln in partial.splitlines() if ln.strip()][:max_lines] + [f"timed out after {timeout} s"]
    except Exception as ex:
        return [repr(ex)[:300]]


def experiments_sample(args, main_checksum, K, W):
    """Secondary, N = 1 only, each in its own child process after every timed region of the main arm (nothing here can touch the headline):
    (1) the experimental megakernel with the down projection fused into the gate_up phase on the main arm's exact workload — tokens/s and
    whether its tokens equal the main arm's; (2) tools/microbench/fusion_probe — what combining 148 partial vectors with vector reductions
    costs against the grid barrier, cluster co-residency, TMA ingest from L2-resident against HBM-resident tiles;
    (3) the per-phase timelines of both kernels."""
    out = {}
    if args.wdtype in ("f32", "bf16"):
        lines, note = _child_json_lines([sys.executable, os.path.join(ROOT, "tools", "fuse_bench.py"), "--config", args.config, "--wdtype", args.wdtype,
                                         "--kvdtype", args.kvdtype, "--prompt-len", str(args.prompt_len), "--steps", str(K), "--warmup", str(W),
                                         "--expect-checksum", str(main_checksum)], 120)
        out["fused_down"] = lines[-1] if lines else {"error": note or "no output"}
        if lines and note:
            out["fused_down"]["note"] = note
        if out["fused_down"].get("tokens_match_main_arm") is not True:
            # something is off with the never-before-run kernel: bring back WHERE (layout of the transposed matrix alone; one forward of a
            # one-layer model through the verified and the fused kernel, buffer by buffer, for 1 / 2 / 16 stripes; the reference's golden stream)
            out["fused_down_diagnosis"] = _child_text([sys.executable, os.path.join(ROOT, "tests", "fuse_check.py"), "--quick", "/tmp/sllm_fuse_check.log"], 60, max_lines=40)
    probe = os.path.join(ROOT, "tools", "microbench", "_build", "fusion_probe")
    if os.path.exists(probe):
        lines, note = _child_json_lines([probe, "1000"], 60)
        out["fusion_probe"] = lines if lines else {"error": note or "no output"}
    return out


def trace_sample(args):
    """(3) where the time of a decode step goes: per-phase timeline (%globaltimer stamps) of the default megakernel and of the fused one,
    four layers of the same widths (tools/mega_trace.py). Last in line for the shared budget: the longest-running child (the batched
    decoder) goes before it, the two short experiments before that."""
    if not (args.config == "llama2-7b" and args.wdtype == "bf16" and args.kvdtype == "bf16"):
        return None
    trace = [sys.executable, os.path.join(ROOT, "tools", "mega_trace.py"), "--pos", str(args.prompt_len)]
    return {"megakernel": _child_text(trace, 60), "megakernel(fused-down)": _child_text(trace + ["--fuse-down"], 60)}


def golden_token_check(args, tokens, pos_first):
    """The tokens of the timed steps against the CPU oracle's greedy stream on this exact workload (tests/golden/bench_cfg4_stream.npz,
    produced by tests/golden/make_bench_golden.py: 32 layers, bf16-rounded weights and cache rows, the same 512-token prompt).
    Reported, not asserted: at 32 layers the synthetic gain-4 model amplifies fp32 rounding noise chaotically — the reference
    algorithm built with FMA contraction leaves its own strict build's stream within a handful of tokens (tests/test_full_config_gpu.py
    has the measurement and the per-position logit check that replaces token identity at this depth)."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "bench_cfg4_stream.npz")
    if not (args.config == "llama2-7b" and args.wdtype == "bf16" and args.kvdtype == "bf16" and args.prompt_len == PROMPT_LEN and os.path.exists(path)):
        return None
    g = np.load(path)
    # `tokens` = everything generated since the prompt ended: tokens[0] follows position PROMPT_LEN-1; golden[i] follows position i
    gold = g["tokens"][PROMPT_LEN - 1:PROMPT_LEN - 1 + tokens.size]
    n = min(gold.size, tokens.size)
    diff = tokens[:n] != gold[:n]
    same = int(np.argmax(diff)) if diff.any() else n
    timed0 = pos_first - (PROMPT_LEN - 1)
    return {"against": "tests/golden/bench_cfg4_stream.npz (CPU oracle, full depth)", "generated_compared": n,
            "generated_identical_prefix": same, "timed_steps_identical": int((~diff[timed0:n]).sum()), "timed_steps": int(n - timed0),
            "first_timed_position": int(pos_first), "oracle_first_generated_token": int(g["tokens"][PROMPT_LEN - 1]),
            "oracle_min_margin_generated": float(g["margins"].min()),
            "note": "reported, not asserted: after this 512-token prompt the gain-4 synthetic model is fully chaotic at 32 layers - the reference algorithm "
                    "itself, rebuilt with FMA contraction or -Ofast, moves the logits of the last prompt position by 1.0-1.2 x max|logit| and shares NO generated "
                    "token with its strict build (profiles/r02_chaos_yardstick_bench_prompt.txt, tools/chaos_yardstick.py). Parity at this depth is "
                    "checked per position on identical cache contents instead (tests/test_full_config_gpu.py)"}


# ------------------------------------------------------------------------------------------------ our arm --
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from simplellminference_b200.config import PRESETS, F32, BF16, INT8
    from simplellminference_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ms = PRESETS[args.config]
    wd = {"f32": F32, "bf16": BF16, "int8": INT8}[args.wdtype]
    kvd = {"f32": F32, "bf16": BF16}[args.kvdtype]
    K, W, P = args.steps, max(args.warmup, 3), args.prompt_len
    assert P + W + 2 * K + 8 <= ms.max_len, "prompt + steps exceed the model's max_len"

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    v2 = world == 1 and not (args.mega_v1 or args.mega_ll or args.no_mega or args.unfused)
    # fusing the down projection into the gate_up phase trades a grid barrier for 148 partial vectors added in L2: it pays when the down matrix is
    # large (Llama-2-7B 90 MB: +7 %, Llama-3-8B 117 MB: +6 %) and costs on small ones (TinyLlama 23 MB: -4 %; profiles/r02_other_configs.md)
    fuse_pays = ms.inter * ms.hidden * {"f32": 4, "bf16": 2, "int8": 1}[args.wdtype] >= 48e6
    eng = Engine(ms, w_dtype=wd, kv_dtype=kvd, tp_rank=rank, tp_size=world, stream=stream, fused=not args.unfused,
                 graph=not (args.no_graph or args.unfused), pdl=args.pdl, mega=not (args.no_mega or args.unfused), mega_ll=args.mega_ll,
                 p2p_allreduce=(world > 1 and not args.nccl), mega_fuse_down=((args.mega_fuse_down or (v2 and fuse_pays)) and world == 1), mega_v2=v2)
    eng.load_synthetic(1234)
    calibrated = False
    if world == 1 and not args.no_calibrate and eng.mode.startswith("megakernel") and "ll" not in eng.mode:
        eng.calibrate(3)   # part of engine set-up (like building a CUDA graph): 27 untimed steps from position 0, then the state is reset
        calibrated = True
    if world > 1:
        eng.init_comm(dist) if args.nccl else eng.init_p2p(dist)
        if not args.nccl and eng.prefill_supported and not getattr(eng, "prefill_p2p", False):
            eng.init_comm(dist)   # no prefill exchange block: batched prefill all-reduces its [T][d] partial sums with NCCL

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms_val):
        if world == 1:
            return ms_val
        t = torch.tensor([ms_val], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.kernel_only:   # profiling aid (ncu): a short loop of one fused kernel over the layers
        eng.set_state(1, P)
        for it in range(3):
            for l in range(ms.layers):
                eng.enqueue_kernel(args.kernel_only, l)
        torch.cuda.synchronize()
        print(json.dumps({"kernel_only": args.kernel_only, "launches": 3 * ms.layers}))
        return

    ids = prompt_ids(P, ms.vocab)
    # ---- batched prefill of the same prompt on the tensor cores (tcgen05 GEMMs), timed on its own: prompt tokens/s and
    # the fraction of the measured dense-bf16 peak. Host ids -> device inside the timed region (it is the public call).
    prefill = None
    if eng.prefill_supported:
        for _ in range(2):
            eng.prefill(ids)
        barrier()
        R = 5
        evp = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
        evp[0].record(stream)
        for r in range(R):
            eng.prefill(ids)
            evp[r + 1].record(stream)
        barrier()
        pf_ms = max_over_ranks(float(np.median([evp[r].elapsed_time(evp[r + 1]) for r in range(R)])))
        pf_noex_ms = pf_trace = None
        if world > 1 and getattr(eng, "prefill_p2p", False):
            torch.cuda.synchronize()
            tr = eng.buffer(201).view(torch.int64).cpu().numpy()   # stamps of the last full exchange (all rows) and of the final-norm one (one row)
            def _us(t):
                return {"launch_to_gemm_done": (t[1] - t[0]) / 1e3, "entry_flags": (t[2] - t[1]) / 1e3, "rows": (t[3] - t[2]) / 1e3,
                        "fence": (t[4] - t[3]) / 1e3, "exit_flags_incl_other_ctas": (t[5] - t[4]) / 1e3}
            pf_trace = {"what": "CTA 0 of rank 0, %globaltimer", "full_call_us": _us(tr[0:8]), "one_row_call_us": _us(tr[8:16])}
            from simplellminference_b200 import _lib
            _lib.check(eng.lib.sllm_tune(8, 16))
            eng.prefill(ids)
            barrier()
            evq = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
            evq[0].record(stream)
            for r in range(R):
                eng.prefill(ids)
                evq[r + 1].record(stream)
            barrier()
            pf_noex_ms = max_over_ranks(float(np.median([evq[r].elapsed_time(evq[r + 1]) for r in range(R)])))
            _lib.check(eng.lib.sllm_tune(8, 0))
            eng.prefill(ids)   # leave a valid cache / state behind
            barrier()
        tpeak, tsrc = tensor_peak()
        tf = prefill_flops(ms, P) / (pf_ms * 1e-3) / 1e12
        prefill = {"tokens": P, "ms": pf_ms, "tokens_per_sec": P / (pf_ms * 1e-3), "tflops": tf, "peak_tflops": tpeak * world,
                   "tensor_pipe_frac": tf / (tpeak * world), "peak_source": tsrc, "reps": R,
                   "what": "sllm_engine_prefill: all layers over the whole prompt (tcgen05/TMEM GEMMs + block attention) + last-row logits/arg-max; "
                           "algorithmic flops per SURVEY.md 8d",
                   "ms_without_exchanges": pf_noex_ms, "exchange_trace": pf_trace,
                   "tp_exchange": None if world == 1 else ("peer memory (csrc/prefill_tp.cu: sum over ranks + residual + RMSNorm + row distribution in one kernel)"
                                                           if getattr(eng, "prefill_p2p", False) else "ncclAllReduce + RMSNorm kernel")}
    # ---- prompt for the decode measurement: fed token by token through the decode step exactly as the reference does
    # (model.cpp:157-166), untimed; it leaves the KV cache filled for positions 0..P-1 with fp32-activation values
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_prompt0 = time.perf_counter()
    toks = eng.greedy(ids, P + 1)            # P forwards: positions 0..P-1; state is now (argmax, pos=P)
    t_prompt = time.perf_counter() - t_prompt0
    assert toks.size == P and np.array_equal(toks[:P - 1], ids[1:]), "prompt echo mismatch"

    # ---- resident decode: W warm-up + K timed graph replays, token feedback on the device
    eng.enqueue_steps(W)
    barrier()
    sampler.mark(True)
    launches0 = eng.total_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    eng.enqueue_steps(K)
    ev1.record(stream)
    barrier()
    sampler.mark(False)
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.total_launches - launches0
    pos_first = P + W
    generated = eng.read_tokens(W + K + 1)   # the first generated token (output of position P-1), the warm-up steps, the timed steps
    tokens = generated[-K:]
    value = K / (ms_total * 1e-3)
    step_bytes = float(np.mean([eng.step_bytes(p) for p in range(pos_first, pos_first + K)]))
    full_bytes = float(np.mean([ms.step_bytes(p, wd, kvd) for p in range(pos_first, pos_first + K)]))

    # ---- e2e: the reference-facing call, host buffers, one sync per token
    pos = pos_first + K
    tok = int(tokens[-1])
    for _ in range(3):
        _, tok = eng.forward(tok, pos, want_logits=False); pos += 1
    barrier()
    Ke = min(K, ms.max_len - pos - 1)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for _ in range(Ke):
        _, tok = eng.forward(tok, pos, want_logits=False); pos += 1
    ev3.record(stream)
    barrier()
    e2e_ms = max_over_ranks(ev2.elapsed_time(ev3))
    e2e = {"value": Ke / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 4,
           "steps": Ke, "api": "Engine.forward(token,pos) == sllm_engine_forward: LlamaModel::forward semantics + device argmax"}

    # ---- roofline of the dominant kernel. Megakernel mode: the step IS one kernel (100 % of the timed region), so
    # achieved = B(p) / its launch duration over the timed region. Per-kernel mode: the largest kernel (gate_up:
    # RMSNorm + [Wup;Wgate] GEMV + sigmoid*up) timed alone, cycling over the layers so weights come from HBM.
    roof = None
    peak, peak_src = peaks()
    mode = eng.mode
    if mode.startswith("megakernel"):
        ach = step_bytes / (ms_total * 1e-3 / K) / 1e9
        # DRAM bytes of one launch from the committed ncu --set full capture of the same kernel, configuration and position range
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "mega_traffic.json")) as f:
                tj = json.load(f)
            if world == 1 and args.config == "llama2-7b" and args.wdtype == "bf16" and args.kvdtype == "bf16" and mode in tj:
                traffic, traffic_src = tj[mode]["traffic_bytes_per_launch"], tj[mode]["source"]
        except Exception:
            traffic = None
        kname = {"megakernel": "mega_step_kernel (persistent: all layers' qkv|attention|wo|gate_up|down + classifier/argmax, 5 grid barriers per layer)",
                 "megakernel(v2,fused-down)": "mega2_step_kernel (persistent: all layers' qkv->attention->wo | gate_up->down + classifier/argmax, 2 grid barriers per layer)"}.get(mode, mode)
        roof = {"bound": "hbm", "kernel": kname,
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "bytes_per_launch": step_bytes,
                "us_per_launch": 1e3 * ms_total / K, "peak_source": peak_src, "traffic_source": traffic_src,
                "how": f"{K} launches = the timed region itself, CUDA events on the launching stream; B(p) per SURVEY.md 8d"}
    elif not args.unfused:
        eng.set_state(tok, min(pos, ms.max_len - 1))
        reps = max(2, 64 // ms.layers)
        for l in range(ms.layers):
            eng.enqueue_kernel("gate_up", l)
        barrier()
        ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev4.record(stream)
        for _ in range(reps):
            for l in range(ms.layers):
                eng.enqueue_kernel("gate_up", l)
        ev5.record(stream)
        barrier()
        k_ms = max_over_ranks(ev4.elapsed_time(ev5)) / (reps * ms.layers)
        kb = eng.kernel_bytes("gate_up", 0)
        ach = kb / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "fused_gemv_kernel<GateUpPolicy> (RMSNorm + up/gate GEMV + sigmoid(gate)*up)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "bytes_per_launch": kb, "us_per_launch": 1e3 * k_ms, "peak_source": peak_src,
                "how": f"{reps * ms.layers} back-to-back launches cycling over {ms.layers} layers (weights {kb * ms.layers / 1e9:.1f} GB >> L2), CUDA events"}
    clocks = sampler.stop() if rank == 0 else None
    token_check = golden_token_check(args, generated, pos_first) if rank == 0 else None

    ach_step = step_bytes * world / (ms_total * 1e-3 / K) / 1e9   # aggregate over ranks
    step_roof = {"bound": "hbm", "achieved": ach_step, "peak": peak * world, "unit": "GB/s", "frac": ach_step / (peak * world),
                 "frac_of_8TBs_nominal": ach_step / (8000.0 * world), "bytes_per_step": full_bytes, "peak_source": peak_src,
                 "positions": [pos_first, pos_first + K - 1]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    step_launches = eng.step_launches
    batch = experiments = None
    if world == 1 and args.extras:
        eng.close()
        try:   # secondary: never lose the line over any of it. Order = cost: the two short experiments, the batched decoder, the timelines
            if mode == "megakernel":   # the round-1 kernel: measure its fused-down variant beside it
                experiments = experiments_sample(args, int(np.sum(tokens.astype(np.int64)) % 1000003), K, W)
            batch = batch_decode_sample(args)
            # the parity cases of the batched decoder that no GPU has run yet (pytest reports them as xfail / xpass only): PASS / FAIL per
            # case with the traceback of a failure — tests/batch_check.py --quick
            batch["never_run_parity_cases"] = _child_text([sys.executable, os.path.join(ROOT, "tests", "batch_check.py"), "--quick", "/tmp/sllm_batch_check.log"], 90, max_lines=60)
            if mode == "megakernel":
                experiments["mega_trace"] = trace_sample(args)
        except Exception as ex:
            experiments = dict(experiments or {}, error=repr(ex)[:300])
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported at N = 1 only (it costs ~10 s of host time)
        try:
            cpu = cpu_reference_sample(ms, P, 3, 1, budget_s=60.0)   # bounded: ~20 s of host time at 7B (3 + 1 full-depth forwards)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is informational; never lose the GPU numbers over it
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(ex)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.wdtype, "data": "synthetic",
        "config": workload_config(args, ms), "roofline": roof, "step_roofline": step_roof, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "launches_per_step": step_launches, "clocks": clocks,
        "prefill": prefill, "prompt_tokens_per_sec_token_by_token": P / t_prompt, "token_checksum": int(np.sum(tokens.astype(np.int64)) % 1000003),
        "token_check": token_check, "first_generated_token": int(toks[P - 1]), "calibrated_partition": calibrated,
        "mode": mode + ("" if world == 1 else (" tp/nccl-allreduce" if args.nccl else " tp/peer-memory-allreduce")),
        "batch_decode": batch, "experiments": experiments,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: anything a library prints there while we run (NCCL's version banner, ...)
    # is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
