"""tools/v2_locate.py — first position at which an engine mode leaves the oracle on stories110M fp32 (teacher-forced on the oracle's own
stream): per-position max|dlogit| / max|logit| for each sllm_tune(8) value given (megakernel2 A/B bits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import loader
from simplellminference_b200 import _lib
from simplellminference_b200.config import PRESETS, F32
from simplellminference_b200.engine import Engine
ms = PRESETS["stories110M"]
port = loader.Port()
sh = loader.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)
blob = port.fill_blob(sh, 1234, loader.F32, 64)
om = port.model(sh, blob, threads=os.cpu_count() or 1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
toks, want = [1], []
for pos in range(N):
    lg = om.forward(toks[-1], pos)
    want.append(lg)
    toks.append(int(np.argmax(lg)))
lib = _lib.load()
for mode, kw in (("v1", dict(mega=True)), ("v2", dict(mega=True, mega_v2=True)), ("v2f", dict(mega=True, mega_v2=True, mega_fuse_down=True))):
    for dbg in ((0,) if mode == "v1" else (0, 4, 8, 12)):
        lib.sllm_tune(8, dbg)
        eng = Engine(ms, w_dtype=F32, kv_dtype=F32, **kw).load_synthetic(1234)
        errs = []
        for pos in range(N):
            got, nxt = eng.forward(toks[pos], pos)
            errs.append(float(np.abs(got - want[pos]).max()) / float(np.abs(want[pos]).max()))
        errs = np.array(errs)
        bad = np.flatnonzero(errs > 3e-4)
        print(f"{mode} {eng.mode} debug={dbg}: max rel err {errs.max():.2e} at {int(errs.argmax())}; first position above 3e-4: {int(bad[0]) if bad.size else None}; "
              f"errs at 120..150 step 6: {[f'{e:.1e}' for e in errs[120:150:6]]}", flush=True)
        eng.close()
lib.sllm_tune(8, 0)
