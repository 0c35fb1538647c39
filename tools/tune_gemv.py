"""tools/tune_gemv.py — development aid: per-kernel HBM throughput of the fused decode kernels at a model shape,
each kernel kind timed alone (cycling over the layers so weights come from HBM) and the full graph step."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplellminference_b200 import _lib
from simplellminference_b200.config import PRESETS, BF16, F32, INT8
from simplellminference_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="llama2-7b")
ap.add_argument("--ctas", default="0")
ap.add_argument("--pos", type=int, default=512)
ap.add_argument("--wdtype", default="bf16")
ap.add_argument("--pdl", action="store_true")
ap.add_argument("--mega", action="store_true")
ap.add_argument("--ll", action="store_true")
a = ap.parse_args()
ms = PRESETS[a.config]
wd = dict(f32=F32, bf16=BF16, int8=INT8)[a.wdtype]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
lib = _lib.load()
for ctas in [int(c) for c in a.ctas.split(",")]:
    lib.sllm_tune(0, ctas)
    eng = Engine(ms, w_dtype=wd, kv_dtype=BF16, stream=stream, pdl=a.pdl, mega=a.mega, mega_ll=a.ll).load_synthetic(1)
    eng.set_state(1, a.pos)
    res = {"ctas_per_sm": ctas, "lib": os.path.basename(_lib.LIB_PATH), "mode": eng.mode}
    for kind in ([] if a.mega else ("qkv", "mha", "wo", "gate_up", "down")):
        for l in range(ms.layers): eng.enqueue_kernel(kind, l)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            for l in range(ms.layers): eng.enqueue_kernel(kind, l)
        e1.record(stream); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * ms.layers)
        res[kind] = {"us": round(us, 2), "GBps": round(eng.kernel_bytes(kind, a.pos) / us / 1e3, 0)}
    eng.set_state(1, a.pos); eng.enqueue_steps(4); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); eng.enqueue_steps(32); e1.record(stream); torch.cuda.synchronize()
    res["step_ms"] = round(e0.elapsed_time(e1) / 32, 4)
    res["tok_s"] = round(1e3 / res["step_ms"], 1)
    if not a.mega: res["sum_kernels_ms"] = round(sum(res[k]["us"] for k in ("qkv", "mha", "wo", "gate_up", "down")) * ms.layers / 1e3, 4)
    res["step_GBps"] = round(eng.step_bytes(a.pos) / res["step_ms"] / 1e6, 0)
    print(json.dumps(res), flush=True)
    eng.close()
