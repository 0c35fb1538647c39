"""tools/mega_sweep.py — the decode megakernel variants side by side on the bench workload's shape and position: ms per token, tokens/s,
fraction of the HBM roofline, checksum of the produced tokens. One JSON line per variant.
(profiles/r02_l2_lookahead_sweep.jsonl was produced by an earlier form of this script that also swept the since-removed L2 look-ahead.)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simplellminference_b200 import _lib
from simplellminference_b200.config import PRESETS, BF16, F32, INT8
from simplellminference_b200.engine import Engine
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="llama2-7b"); ap.add_argument("--pos", type=int, default=512); ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--wdtype", default="bf16"); ap.add_argument("--kvdtype", default="bf16")
ap.add_argument("--variants", default="v1,v2f,v2f+cal"); ap.add_argument("--debug", default="0", help="comma list of sllm_tune(8) values (megakernel2 A/B bits: 4 = scalar wo reductions, 8 = one x replica)")
a = ap.parse_args()
ms = PRESETS[a.config]
WD = {"f32": F32, "bf16": BF16, "int8": INT8}[a.wdtype]; KVD = {"f32": F32, "bf16": BF16}[a.kvdtype]
lib = _lib.load()
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
KW = {"v1": {}, "v1f": dict(mega_fuse_down=True), "v2": dict(mega_v2=True), "v2f": dict(mega_v2=True, mega_fuse_down=True), "ll": dict(mega_ll=True)}
for var in a.variants.split(","):
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    eng = Engine(ms, w_dtype=WD, kv_dtype=KVD, stream=stream, mega=True, **KW[var.split("+")[0]]).load_synthetic(1234)
    if var.endswith("+cal"):
        eng.calibrate(3)
        tau = eng.calibration()
        print(json.dumps({"variant": var, "calibration": {"min": float(tau.min()), "max": float(tau.max()), "std": float(tau.std())}}), flush=True)
    for rep, dbg in enumerate([int(x, 0) for x in a.debug.split(",")] if var.startswith("v2") else [0]):
        lib.sllm_tune(8, dbg)
        eng.set_state(1, a.pos); eng.enqueue_steps(5); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); eng.enqueue_steps(a.steps); e1.record(stream); torch.cuda.synchronize()
        msz = e0.elapsed_time(e1) / a.steps
        toks = eng.read_tokens(a.steps + 5)
        nbytes = sum(eng.step_bytes(p) for p in range(a.pos + 5, a.pos + 5 + a.steps)) / a.steps
        print(json.dumps({"variant": var, "mode": eng.mode, "debug_bits": dbg, "ms_per_step": round(msz, 4),
                          "tokens_per_sec": round(1e3 / msz, 1), "gbs": round(nbytes / msz / 1e6), "frac_of_measured_peak": round(nbytes / msz / 1e6 / peak, 3),
                          "token_checksum": int(np.sum(toks.astype(np.int64)) % 1000003)}), flush=True)
    lib.sllm_tune(8, 0)
    eng.close(); del eng
