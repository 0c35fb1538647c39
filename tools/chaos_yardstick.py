"""tools/chaos_yardstick.py — how far does the REFERENCE ALGORITHM move when only its floating-point summation details change,
on bench.py's exact workload (Llama-2-7B shape, 32 layers, bf16-rounded weights and cache rows, the 512-token prompt)?

Three builds of the same C restatement (oracle/Makefile): strict (-O2 -ffp-contract=off, pinned bit for bit to the reference),
FMA (-O3 -march=x86-64-v3 -ffp-contract=fast) and -Ofast (re-associated sums). Each feeds the prompt token by token into ITS OWN
cache, then decodes greedily. Printed per build: its first generated tokens, max|logit difference| to the strict build at the last
prompt position, and where its stream leaves the strict one. bench.py's "token_check" is read against these numbers: a CUDA path
cannot be expected to stay closer to the strict build than the reference's own alternative builds do.   CPU only; ~15 min per build.

    python tools/chaos_yardstick.py [n_generated=8] > profiles/r02_chaos_yardstick_bench_prompt.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from bench import prompt_ids, PROMPT_LEN  # noqa: E402
from simplellminference_b200.config import PRESETS  # noqa: E402


def main():
    n_gen = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ms = PRESETS["llama2-7b"]
    shape = loader.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)
    loader.build("port")
    strict = loader.Port()
    blob = strict.fill_blob(shape, 1234, loader.BF16)
    ids = prompt_ids(PROMPT_LEN, ms.vocab)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bench_cfg4_stream.npz"))
    results = {}
    for name, path in (("strict", loader.PORT_SO), ("fma", loader.PORT_FMA_SO), ("ofast", loader.PORT_FAST_SO)):
        port = loader.Port(path)
        m = port.model(shape, blob, threads=os.cpu_count() or 1, kv_bf16=True)
        t0 = time.time()
        tok, toks, at_prompt_end = int(ids[0]), [], None
        for pos in range(PROMPT_LEN + n_gen - 1):
            if pos < PROMPT_LEN - 1:
                m.step(tok, pos)
                tok = int(ids[pos + 1])
            else:
                logits = m.forward(tok, pos)
                if pos == PROMPT_LEN - 1:
                    at_prompt_end = logits.copy()
                tok = int(np.argmax(logits))
                toks.append(tok)
        m.close()
        results[name] = (toks, at_prompt_end)
        print(f"{name:6s} ({os.path.basename(path)}): generated {toks}   [{time.time() - t0:.0f} s]", flush=True)
    s_toks, s_log = results["strict"]
    want = gold["tokens"][PROMPT_LEN - 1:PROMPT_LEN - 1 + n_gen].tolist()
    print(f"golden stream (tests/golden/bench_cfg4_stream.npz): {want}; strict build reproduces it: {s_toks == want}")
    srt = np.sort(s_log)
    print(f"strict build at the last prompt position: max|logit| {np.abs(s_log).max():.1f}, top-1/top-2 margin {srt[-1] - srt[-2]:.3f}")
    for name in ("fma", "ofast"):
        toks, lg = results[name]
        same = next((i for i, (a, b) in enumerate(zip(toks, s_toks)) if a != b), len(toks))
        print(f"{name:6s} vs strict: max|dlogit| at the last prompt position {np.abs(lg - s_log).max():.3f} "
              f"({np.abs(lg - s_log).max() / np.abs(s_log).max():.2e} of max|logit|), identical generated prefix {same} of {len(toks)}")


if __name__ == "__main__":
    main()
