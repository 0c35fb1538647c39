"""tools/profile_step.py — tiny driver for ncu: a few decode steps of one engine mode at a (possibly layer-cut) shape."""
import argparse, dataclasses, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplellminference_b200.config import PRESETS, BF16, F32, INT8
from simplellminference_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="llama2-7b")
ap.add_argument("--layers", type=int, default=0)
ap.add_argument("--mode", default="mega", choices=["mega", "fused"])
ap.add_argument("--pos", type=int, default=512)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--fuse-down", action="store_true", help="mega mode: the experimental kernel with the down projection fused into the gate_up phase")
ap.add_argument("--v2", action="store_true", help="mega mode: csrc/megakernel2.cu (two grid-wide dependency points per layer)")
ap.add_argument("--calibrate", action="store_true", help="mega mode: sllm_engine_calibrate before the steps")
ap.add_argument("--batch", type=int, default=0, help="> 0: that many sequences through the batched decoder (sllm_batch_*) instead of the engine's own step")
a = ap.parse_args()
ms = PRESETS[a.config]
if a.layers:
    ms = dataclasses.replace(ms, layers=a.layers)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
if a.batch:   # every sequence decodes a.pos tokens first (untimed), then a.steps timed steps
    from simplellminference_b200.batch import BatchDecoder
    ms = dataclasses.replace(ms, max_len=a.pos + a.steps + 8)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, stream=stream).load_synthetic(1)
    bd = BatchDecoder(eng, max_seqs=a.batch, page_len=64, kv_dtype=BF16)
    for i in range(a.batch):
        bd.add([1 + i])
    bd.step(a.pos); torch.cuda.synchronize()
    nbytes = bd.step_bytes()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); bd.step(a.steps); e1.record(stream); torch.cuda.synchronize()
    msec = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"mode": f"batch of {a.batch}", "layers": ms.layers, "step_ms": round(msec, 4), "tokens_per_sec": round(a.batch / msec * 1e3, 1),
                      "GBps": round(nbytes / msec / 1e6, 0)}))
    sys.exit(0)
eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, stream=stream, mega=(a.mode == "mega"), mega_fuse_down=a.fuse_down, mega_v2=a.v2).load_synthetic(1)
if a.calibrate:
    eng.calibrate(3)
eng.set_state(1, a.pos)
eng.enqueue_steps(2); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); eng.enqueue_steps(a.steps); e1.record(stream); torch.cuda.synchronize()
msec = e0.elapsed_time(e1) / a.steps
print(json.dumps({"mode": eng.mode, "layers": ms.layers, "step_ms": round(msec, 4), "GBps": round(eng.step_bytes(a.pos) / msec / 1e6, 0)}))
