"""tests/golden/make_bench_golden.py — the CPU oracle's greedy stream on bench.py's EXACT workload.

    python tests/golden/make_bench_golden.py [n_generated=160]

Llama-2-7B-shaped model at full depth (32 layers), synthetic weights seed 1234 rounded to bf16, bf16-rounded KV rows
(orc_set_kv_bf16 — the definition of the bf16 cache, DESIGN.md "Tolerances"), the 512-token prompt of bench.prompt_ids fed
token by token exactly as LlamaModel::predict does (model.cpp:157-166), then greedy arg-max feedback. Needs ~30 GiB of host
RAM and ~10-15 minutes on 8 cores (the C restatement, pinned bit-for-bit to oracle/_ref by tests/test_oracle_cpu.py, with its
row-parallel matmul threads; the sums inside a row are the reference's serial sums, so threading changes no bit).

Output: tests/golden/bench_cfg4_stream.npz = {tokens (everything after prompt[0]), margins (top1-top2 of every GENERATED
position), last_logits}. bench.py compares the tokens its timed steps produced against this stream ("token_check").
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from bench import prompt_ids, PROMPT_LEN  # noqa: E402
from simplellminference_b200.config import PRESETS  # noqa: E402


def main():
    n_gen = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    ms = PRESETS["llama2-7b"]
    shape = loader.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)
    port = loader.Port()
    t0 = time.time()
    blob = port.fill_blob(shape, 1234, loader.BF16)
    print(f"blob {blob.nbytes / 2**30:.1f} GiB in {time.time() - t0:.0f} s", flush=True)
    m = port.model(shape, blob, threads=os.cpu_count() or 1, kv_bf16=True)
    ids = prompt_ids(PROMPT_LEN, ms.vocab)
    toks, margins = [], []
    tok = int(ids[0])
    t0 = time.time()
    for pos in range(PROMPT_LEN + n_gen - 1):
        logits = m.forward(tok, pos)
        if pos < PROMPT_LEN - 1:
            tok = int(ids[pos + 1])
        else:
            tok = int(np.argmax(logits))
            srt = np.partition(logits, -2)[-2:]
            margins.append(float(srt[1] - srt[0]))
        toks.append(tok)
        if pos % 50 == 0:
            print(f"pos {pos} {time.time() - t0:.0f} s", flush=True)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_cfg4_stream.npz")
    np.savez_compressed(out, tokens=np.array(toks, np.int32), margins=np.array(margins, np.float32), last_logits=logits,
                        prompt_len=np.int32(PROMPT_LEN))
    print(f"wrote {out}: {len(toks)} tokens, min margin over the generated part {min(margins):.4f}", flush=True)


if __name__ == "__main__":
    main()
