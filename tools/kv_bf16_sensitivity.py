import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import loader
from simplellminference_b200.config import ModelShape, BF16, F32
from simplellminference_b200.engine import Engine
os.chdir('/tmp')
ms = ModelShape(4096, 128, 1024, 512, 2816, 160, 4, 8, 4)
shape = loader.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)
port = loader.Port()
blob = port.fill_blob(shape, 1234, BF16, 64)
prompt = list(range(1, 17))
for n_total in (20, 40, 80, 120):
    want, want_l = port.model(shape, blob, threads=8, kv_bf16=True).greedy(prompt, n_total)
    w32, w32_l = port.model(shape, blob, threads=8, kv_bf16=False).greedy(prompt, n_total)
    for mega in (False, True):
        eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, mega=mega).load_synthetic(1234)
        got = eng.greedy([prompt[0]] + [int(t) for t in want[:-1]], n_total)
        lg = eng.buffer("model_pred").cpu().numpy()
        print(n_total, 'mega' if mega else 'fused', 'err vs bf16kv-oracle', float(np.abs(lg - want_l).max()), 'scale', float(np.abs(want_l).max()),
              '| oracle bf16kv vs f32kv', float(np.abs(want_l - w32_l).max()) if np.array_equal(want, w32) else 'tokens differ', flush=True)
        eng.close()
