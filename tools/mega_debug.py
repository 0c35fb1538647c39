"""tools/mega_debug.py — what bounds the decode megakernel: the same step with its grid barriers removed (pure streaming + arithmetic, no
dependency stalls) and with the dot products removed (pure TMA ring streaming). Results of those runs are garbage by construction
(sllm_tune key 8); only the step time is read. One JSON line per variant."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplellminference_b200 import _lib
from simplellminference_b200.config import PRESETS, BF16
from simplellminference_b200.engine import Engine
ap = argparse.ArgumentParser(); ap.add_argument("--config", default="llama2-7b"); ap.add_argument("--pos", type=int, default=512); ap.add_argument("--steps", type=int, default=20); ap.add_argument("--v2", action="store_true"); ap.add_argument("--fuse-down", action="store_true")
a = ap.parse_args()
ms = PRESETS[a.config]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, stream=stream, mega=True, mega_v2=a.v2, mega_fuse_down=a.fuse_down).load_synthetic(1234)
print(json.dumps({"mode": eng.mode}), flush=True)
lib = _lib.load()
for dbg, what in ((0, "normal"), (1, "no grid barriers"), (2, "no dot products"), (3, "no barriers, no dot products"), (0, "normal again")):
    lib.sllm_tune(8, dbg)
    eng.set_state(1, a.pos); eng.enqueue_steps(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); eng.enqueue_steps(a.steps); e1.record(stream); torch.cuda.synchronize()
    msz = e0.elapsed_time(e1) / a.steps
    nbytes = sum(eng.step_bytes(p) for p in range(a.pos + 5, a.pos + 5 + a.steps)) / a.steps
    print(json.dumps({"variant": what, "debug": dbg, "ms_per_step": msz, "tokens_per_sec": 1e3 / msz, "gbs": nbytes / msz / 1e6}), flush=True)
lib.sllm_tune(8, 0)
