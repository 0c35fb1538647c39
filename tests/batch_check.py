"""Stand-alone runner of the batched-decode GPU checks (tests/test_zz_batch_gpu.py) with a progressive log, for a GPU
box on a short lease: `python tests/batch_check.py [logfile]`. First a one-step localisation (every per-slot buffer
of the batch against the per-kernel engine at position 0, so a wrong kernel is named), then each test function in
turn (`--quick`: only the never-run cases that take seconds). Not collected by pytest (no test_ prefix); the assertions live in the
test module."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
QUICK = "--quick" in sys.argv      # only the cases that have never run on a GPU and finish in seconds (bench.py's diagnostic child)
_args = [a for a in sys.argv[1:] if a != "--quick"]
LOG = _args[0] if _args else os.path.join(ROOT, "gpurun_out", "batch_check.log")
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
    from simplellminference_b200.batch import BatchDecoder
    from simplellminference_b200.config import F32, PRESETS
    from simplellminference_b200.engine import Engine
    os.chdir("/tmp")

    # ---- 1. localisation: one step of three slots at position 0 against the engine's own kernels
    try:
        if QUICK:
            raise StopIteration
        ms = PRESETS["tiny_gqa"]
        eng = Engine(ms, w_dtype=F32, kv_dtype=F32).load_synthetic(1234)
        bd = BatchDecoder(eng, max_seqs=4, page_len=4, kv_dtype=F32)
        toks = [1, 5, 9]
        slots = [bd.add([t]) for t in toks]
        bd.step(1)
        log(f"one step enqueued, slots={slots}, launches={bd.total_launches}")
        names = ["query", "mha_output", "ffn_input", "swi_output", "emb_output", "model_pred"]
        got = {n: bd.buffer(n).cpu().numpy() for n in names}
        nxt_b = [int(bd.tokens(s)[0]) for s in slots]
        for s, t in zip(slots, toks):
            _, nxt = eng.forward(t, 0)
            parts = []
            for n in names:
                e = eng.buffer(n).cpu().numpy()
                g = got[n][s][:e.size]
                parts.append(f"{n}={float(np.abs(g - e[:g.size]).max()):.2e}")
            log(f"slot {s} token {t}: next batch/engine {nxt_b[s]}/{nxt}  max|diff| " + " ".join(parts))
        bd.close(); eng.close()
    except StopIteration:
        pass
    except Exception:
        log("localisation FAILED:\n" + traceback.format_exc())

    # ---- 2. the test functions
    import test_zz_batch_gpu as T
    from simplellminference_b200.config import BF16, INT8
    cases = [("ragged f32/f32", lambda: T.test_ragged_batch_matches_oracle_per_sequence(port, F32, F32)),
             ("retire/readmit", lambda: T.test_retire_and_readmit_recycles_slots_and_pages(port)),
             ("ragged bf16/bf16", lambda: T.test_ragged_batch_matches_oracle_per_sequence(port, BF16, BF16)),
             ("hd48 page 1", lambda: T.test_head_dim_48_and_page_of_one_position(port)),
             ("11 sequences", lambda: T.test_more_sequences_than_one_launch_holds(port)),
             ("ragged int8/f32", lambda: T.test_ragged_batch_matches_oracle_per_sequence(port, INT8, F32)),
             ("batch of one", lambda: T.test_batch_of_one_equals_the_engine(port)),
             ("errors", lambda: T.test_argument_and_state_errors(port)),
             ("step bytes", lambda: T.test_step_bytes_share_the_weights(port)),
             ("ragged bf16/f32", lambda: T.test_ragged_batch_matches_oracle_per_sequence(port, BF16, F32)),
             # written after the first GPU run of this script
             ("continuous batching", lambda: T.test_continuous_batching_matches_oracle_per_request(port)),
             ("graph replay", lambda: T.test_graph_replay_matches_direct_launches()),
             ("four-row gemv f32", lambda: T.test_four_row_gemv_body_is_bit_identical("tiny_gqa", F32)),
             ("four-row gemv bf16", lambda: T.test_four_row_gemv_body_is_bit_identical("tiny_gqa", BF16)),
             ("four-row gemv int8", lambda: T.test_four_row_gemv_body_is_bit_identical("tiny_gqa", INT8)),
             ("four-row gemv hd48", lambda: T.test_four_row_gemv_body_is_bit_identical("tiny_mha_hd48", F32)),
             ("full width 7B x 2 layers, 8 sequences", lambda: T.test_full_width_llama2_7b_two_layers_batch_of_eight(port)),
             ("long context 8 kv heads", lambda: T.test_long_context_many_slots_split_kv(port, 8, 8)),
             ("long context gqa", lambda: T.test_long_context_many_slots_split_kv(port, 8, 2)),
             ("per-slot sampling", lambda: T.test_sampling_per_slot_matches_the_single_sequence_sampler(port)),
             ("split down f32", lambda: T.test_split_down_projection_matches_oracle(port, F32)),
             ("split down bf16", lambda: T.test_split_down_projection_matches_oracle(port, BF16)),
             ("split down int8", lambda: T.test_split_down_projection_matches_oracle(port, INT8))]
    import numpy as _np
    gm = dict(_np.load(os.path.join(ROOT, "tests", "golden", "models_ref.npz")))
    for gname in ("cfg1_stories15M", "cfg2_stories110M", "tiny_gqa", "tiny_gqa_bf16w", "tiny_gqa_int8w", "tiny_mha_hd48"):
        cases.append((f"golden stream {gname}", lambda gname=gname: T.test_golden_streams_of_the_reference_inside_a_batch(gm, gname)))
    if QUICK:   # drop what a B200 has already confirmed (the first ten) and the two oracle-heavy shapes
        cases = [c for c in cases[10:] if not c[0].startswith(("full width", "long context"))]
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
