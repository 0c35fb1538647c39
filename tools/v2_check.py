"""tools/v2_check.py — where does the v2 megakernel leave the verified one? Same weights, same (token, position) sequence through both,
buffer by buffer (q, h of the last layer, sigmoid(gate)*up, final residual stream, logits), for several (layers, max_len) cuts of the
Llama-2-7B widths: max_len 64 -> one attention split per head, 4096 -> four."""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simplellminference_b200.config import PRESETS, BF16
from simplellminference_b200.engine import Engine
cases = [(1, 64), (2, 64), (1, 4096), (2, 4096)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for layers, S in cases:
    ms = dataclasses.replace(PRESETS["llama2-7b"], layers=layers, max_len=S)
    for fuse in (False, True):
        a = Engine(ms, w_dtype=BF16, kv_dtype=BF16, mega=True).load_synthetic(9)
        b = Engine(ms, w_dtype=BF16, kv_dtype=BF16, mega=True, mega_v2=True, mega_fuse_down=fuse).load_synthetic(9)
        tok = 1
        for pos in range(4):
            la, na = a.forward(tok, pos)
            lb, nb = b.forward(tok, pos)
            row = {}
            for name in ("query", "ffn_input", "swi_output", "emb_output"):
                x, y = a.buffer(name).float().cpu().numpy(), b.buffer(name).float().cpu().numpy()
                row[name] = f"{np.abs(x - y).max() / max(1e-9, np.abs(x).max()):.1e}"
            row["logits"] = f"{np.abs(la - lb).max() / max(1e-9, np.abs(la).max()):.1e}"
            k0a, k0b = a.kv_row("k", layers - 1, pos).float().cpu().numpy(), b.kv_row("k", layers - 1, pos).float().cpu().numpy()
            row["k_last_layer"] = f"{np.abs(k0a - k0b).max() / max(1e-9, np.abs(k0a).max()):.1e}"
            print(f"L={layers} S={S} {b.mode} pos={pos} next {na} / {nb}  rel diffs: {row}", flush=True)
            tok = na
        a.close(); b.close()
