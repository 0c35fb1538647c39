"""Stand-alone runner for the FIRST GPU execution of the experimental fused-down decode megakernel (SLLM_ENGINE_MEGA_FUSE_DOWN) with a
progressive log: `python tests/fuse_check.py [logfile]`. First a localisation — one forward at position 0 of a ONE-layer model through
the verified megakernel and through the fused one on the same weights, buffer by buffer (ffn_input = h after wo, swi_output =
sigma(gate)*up, emb_output = the residual stream after the down projection, model_pred), for each stripe count — so that a wrong
stage is named; then every case of tests/test_zzz_mega_fuse_gpu.py in turn (`--quick`: only those that need no oracle run). Not collected by pytest
(no test_ prefix)."""
import dataclasses
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
QUICK = "--quick" in sys.argv      # localisation + the cases that need no oracle run (bench.py's diagnostic child: about 20 s)
_args = [a for a in sys.argv[1:] if a != "--quick"]
LOG = _args[0] if _args else os.path.join(ROOT, "gpurun_out", "fuse_check.log")
os.makedirs(os.path.dirname(LOG), exist_ok=True)
_t0 = time.time()
_f = open(LOG, "w")


def log(msg):
    line = f"[{time.time() - _t0:6.2f}s] {msg}"
    print(line, flush=True)
    _f.write(line + "\n")
    _f.flush()
    os.fsync(_f.fileno())


def main():
    import numpy as np
    log("start")
    import torch
    log(f"torch imported, cuda={torch.cuda.is_available()}")
    from oracle import loader
    loader.build("port")
    port = loader.Port()
    from simplellminference_b200.config import BF16, F32, PRESETS, ModelShape
    from simplellminference_b200.engine import Engine
    os.chdir("/tmp")

    # ---- 1. localisation
    shapes = [("tiny_gqa f32 (1 stripe)", dataclasses.replace(PRESETS["tiny_gqa"], layers=1), F32),
              ("d256 bf16 (1 stripe)", ModelShape(1000, 64, 256, 128, 704, 40, 1, 4, 2), BF16),
              ("d512 bf16 (2 stripes)", ModelShape(1000, 64, 512, 512, 1408, 40, 1, 8, 8), BF16),
              ("d4096 bf16 (16 stripes, 7B widths)", dataclasses.replace(PRESETS["llama2-7b"], layers=1, max_len=64), BF16)]
    for name, ms, wd in shapes:
        try:
            kvd = BF16 if ms.head_dim > 64 else F32
            plain = Engine(ms, w_dtype=wd, kv_dtype=kvd, mega=True).load_synthetic(3)
            fused = Engine(ms, w_dtype=wd, kv_dtype=kvd, mega=True, mega_fuse_down=True).load_synthetic(3)
            log(f"{name}: modes {plain.mode} / {fused.mode}")
            lp, np_ = plain.forward(17, 0)
            lf, nf = fused.forward(17, 0)
            parts = []
            for buf in ("ffn_input", "swi_output", "emb_output"):
                a, b = plain.buffer(buf).cpu().numpy(), fused.buffer(buf).cpu().numpy()
                parts.append(f"{buf}={float(np.abs(a - b).max()):.2e}/{float(np.abs(a).max()):.2e}")
            parts.append(f"model_pred={float(np.abs(lp - lf).max()):.2e}/{float(np.abs(lp).max()):.2e}")
            log(f"{name}: next {np_}/{nf}  max|plain - fused| / max|plain|: " + " ".join(parts))
            a = plain.greedy([1, 2, 3], 20)
            b = fused.greedy([1, 2, 3], 20)
            log(f"{name}: 19 greedy tokens {'IDENTICAL' if np.array_equal(a, b) else 'DIFFER at ' + str(int(np.flatnonzero(a != b)[0]))}")
            plain.close(); fused.close()
        except Exception as ex:
            log(f"{name}: localisation FAILED: {type(ex).__name__}: {str(ex)[:300]}" + ("" if QUICK else "\n" + traceback.format_exc(limit=6)))

    # ---- 2. the test functions
    import test_zzz_mega_fuse_gpu as T
    cases = [(f"transposed layout d={d} inter={inter} {dt}", lambda d=d, inter=inter, dt=dt: T.test_transposed_down_layout(d, inter, dt))
             for d, inter, dt in [(128, 384, "f32"), (256, 704, "bf16"), (4096, 11008, "bf16"), (2048, 5632, "f32")]]
    cases += [("golden stream tiny_gqa", T.test_golden_stream_of_the_reference),
             ("fallback shapes", T.test_shapes_it_does_not_take_fall_back_visibly)]
    for d, heads, kvh, inter, wd in ([] if QUICK else [(256, 4, 2, 704, BF16), (512, 8, 8, 1408, BF16), (1024, 16, 4, 2824, BF16), (256, 4, 4, 516, F32)]):
        cases.append((f"oracle d={d} inter={inter} wd={wd}", lambda d=d, heads=heads, kvh=kvh, inter=inter, wd=wd: T.test_stripe_counts_against_the_oracle(port, d, heads, kvh, inter, wd)))
    if not QUICK:
        cases.append(("full width 7B x 2 layers", lambda: T.test_full_width_llama2_7b_two_layers(port)))
    ok = 0
    for name, fn in cases:
        try:
            fn()
            ok += 1
            log(f"PASS {name}")
        except Exception as ex:
            if QUICK:   # one line per failure: the exception and the innermost frame of this repository
                tb = [f for f in traceback.extract_tb(ex.__traceback__) if ROOT in f.filename]
                where = f"{os.path.relpath(tb[-1].filename, ROOT)}:{tb[-1].lineno}" if tb else "?"
                log(f"FAIL {name}: {type(ex).__name__}: {str(ex)[:300]} (at {where})")
            else:
                log(f"FAIL {name}:\n" + traceback.format_exc(limit=6))
    log(f"done: {ok}/{len(cases)} passed")


if __name__ == "__main__":
    main()
