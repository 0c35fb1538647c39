"""tests/golden/make_golden.py — regenerates tests/golden/*.npz from the UNMODIFIED reference CPU path.

Run in the dev container (needs /root/reference):   python tests/golden/make_golden.py
It builds oracle/_ref (oracle/Makefile: g++ -O2 -ffp-contract=off on the reference sources where they lie),
feeds it seeded inputs and stores the reference's outputs. The reference itself ships no golden vectors
(SURVEY.md §4), so these files ARE the pin: tests compare the C restatement (oracle/llama_oracle.c) and the
CUDA path against them on any box, including the GPU box where /root/reference does not exist.
Inputs are not stored: they are regenerated from the seeds below (numpy PCG64 / the synthetic weight hash).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# shapes used by the fixtures (name -> loader.Shape); cfg1/cfg2 are BASELINE.json configs[0]/[1]
SHAPES = {
    "cfg1_stories15M": loader.Shape(32000, 48, 288, 288, 768, 256, 6, 6, 6),
    "cfg2_stories110M": loader.Shape(32000, 64, 768, 768, 2048, 1024, 12, 12, 12),
    "tiny_gqa": loader.Shape(512, 32, 128, 64, 384, 48, 3, 4, 2),
    "tiny_mha_hd48": loader.Shape(300, 48, 96, 96, 200, 40, 2, 2, 2),
}
MODEL_RUNS = {  # name -> (prompt, n_total, wdtype, positions whose full logits are stored)
    "cfg1_stories15M": ([1], 128, loader.F32),
    "cfg2_stories110M": ([1], 24, loader.F32),
    "cfg2_stories110M_256": ([1], 256, loader.F32),          # BASELINE.json configs[1] at its full length: 256 greedy tokens
    "tiny_gqa": ([1, 7, 300, 12, 44], 46, loader.F32),       # S=48, H/KVH=2 -> parity domain pos <= 46
    "tiny_gqa_bf16w": ([1, 7, 300, 12, 44], 46, loader.BF16),
    "tiny_gqa_int8w": ([1, 7, 300, 12, 44], 46, loader.INT8),
    "tiny_mha_hd48": ([5], 40, loader.F32),
}
SEED = 1234


def op_inputs(seed=7):
    """Seeded op-level inputs shared by make_golden.py and the tests."""
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    d, I, hd, H, KVH, S, L = 288, 768, 48, 6, 2, 64, 2
    logits = np.round(f(4096) * 4) / 4
    logits[[3000, 411, 2047]] = logits.max() + 0.25   # a three-way tie for the maximum: first one must win
    return dict(d=d, I=I, hd=hd, H=H, KVH=KVH, S=S, L=L, eps=1e-5, theta=10000.0, pos=41, layer=1,
                x=f(d), w=1.0 + 0.02 * f(d), W=f(77, d) * 0.2, a=f(d), b=f(d), up=f(I), gate=3.0 * f(I),
                q=f(H * hd), k=f(H * hd), kc=f(L, S, KVH * hd), vc=f(L, S, KVH * hd),
                table=f(100, d), token=37, logits=logits)


def main():
    loader.build("ref")
    loader.build("port")
    ref, port = loader.Ref(), loader.Port()
    assert ref.flags == "g++ -O2 -ffp-contract=off", ref.flags

    i = op_inputs()
    sin_c, cos_c = ref.rope_cache(i["hd"], i["S"], i["theta"])
    rq, rk = ref.rope(i["q"], i["k"], i["pos"], sin_c, cos_c, i["hd"])
    ops = dict(
        rmsnorm=ref.rmsnorm(i["x"], i["w"], i["eps"]), matmul=ref.matmul(i["x"], i["W"]),
        add=ref.add(i["a"], i["b"]), swiglu=ref.swiglu(i["up"], i["gate"]),
        embedding=ref.embedding(i["token"], i["table"]), sin=sin_c, cos=cos_c, rope_q=rq, rope_k=rk,
        mha=ref.mha(i["q"], i["kc"], i["vc"], i["layer"], i["pos"], i["hd"], i["H"], i["KVH"]),
        argmax=np.int32(ref.argmax(i["logits"])),
    )
    np.savez_compressed(os.path.join(OUT, "ops_ref.npz"), **ops)

    models = {}
    for name, (prompt, n_total, wd) in MODEL_RUNS.items():
        shape = SHAPES[name.replace("_bf16w", "").replace("_int8w", "").replace("_256", "")]
        blob = port.fill_blob(shape, SEED, wd, 64)
        m = ref.model(shape, blob)
        toks, last = m.greedy(prompt, n_total)
        models[name + "/tokens"] = toks
        models[name + "/last_logits"] = last
        # K cache row of layer 0 at the last position written, and the residual stream after the last forward
        kv = shape.kv_hidden
        models[name + "/k_l0_last"] = m.read(2, (n_total - 2) * kv, kv)
        models[name + "/x_last"] = m.read(4, 0, shape.hidden)
        srt = np.sort(last)
        print(f"{name:20s} tokens={toks.size} distinct={len(set(toks.tolist()))} last margin={srt[-1]-srt[-2]:.4f}")
        m.close()
    np.savez_compressed(os.path.join(OUT, "models_ref.npz"), **models)
    for fn in ("ops_ref.npz", "models_ref.npz"):
        print(fn, os.path.getsize(os.path.join(OUT, fn)), "bytes")


if __name__ == "__main__":
    main()
