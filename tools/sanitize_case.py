"""tools/sanitize_case.py <case> — one small decode run for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
  compute-sanitizer --tool racecheck python tools/sanitize_case.py v2f_7b2
Cases: <mode>_<shape>; modes v1, v1f, v2, v2f, ll, graph; shapes tiny (tiny_gqa), med (d 1024, GQA 4, hd 128), 7b2 (Llama-2-7B widths, 2 layers,
max_len 64 -> one attention split per head), 7b2s (the same with max_len 4096 -> four splits). Prints the tokens and whether they match the
deterministic grid-barrier kernel."""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from simplellminference_b200.config import PRESETS, BF16, F32, ModelShape
from simplellminference_b200.engine import Engine
case = sys.argv[1] if len(sys.argv) > 1 else "v2f_tiny"
mode, shape = case.split("_")
KW = {"v1": dict(mega=True), "v1f": dict(mega=True, mega_fuse_down=True), "v2": dict(mega=True, mega_v2=True), "v2f": dict(mega=True, mega_v2=True, mega_fuse_down=True),
      "ll": dict(mega=True, mega_ll=True), "graph": {}}[mode]
ms, wd, kvd, n = {"tiny": (PRESETS["tiny_gqa"], F32, F32, 30), "med": (ModelShape(4096, 128, 1024, 256, 2816, 160, 4, 8, 2), BF16, BF16, 24),
                  "7b2": (dataclasses.replace(PRESETS["llama2-7b"], layers=2, max_len=64), BF16, BF16, 14),
                  "7b2s": (dataclasses.replace(PRESETS["llama2-7b"], layers=2, max_len=4096), BF16, BF16, 14)}[shape]
eng = Engine(ms, w_dtype=wd, kv_dtype=kvd, **KW).load_synthetic(9)
got = eng.greedy([1, 2, 3], n)
print(case, eng.mode, got.tolist(), flush=True)
if os.environ.get("SLLM_COMPARE", "1") == "1":
    ref = Engine(ms, w_dtype=wd, kv_dtype=kvd, mega=True).load_synthetic(9)
    want = ref.greedy([1, 2, 3], n)
    print("matches the grid-barrier megakernel:", bool(np.array_equal(got, want)), want.tolist(), flush=True)
