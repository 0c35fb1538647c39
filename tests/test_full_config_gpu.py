"""GPU parity at the BASELINE.json configurations THEMSELVES (full depth, full width, late positions) — VERDICT r1 "next" item 1.

The small-shape tests cannot reach what decides correctness at depth: the 129-entry phase table, the barrier / counter epochs over
hundreds of dependency points per step, ring run-ahead across 32 layers, split-KV attention past position 512 with 32 KV heads
(cfg4), past 4096 with GQA-4 (cfg5), int8 group-64 tiles at 22 layers near position 2000 (cfg3). A full-depth CPU forward costs
0.3-0.7 s on the box's cores, so two kinds of check are affordable:

  * head:  a short prompt, then TEACHER-FORCED along the oracle's own greedy stream (LlamaModel::predict, model.cpp:148-185): logits
           at every position, arg-max where the oracle's margin allows it, and the length of the identical greedy prefix is reported;
  * late:  a synthetic KV history INJECTED into both sides (the same bf16-exact values in the oracle's [L][S][kv] cache and in the
           engine's cache, whatever its layout), then teacher-forced forwards at the late positions. Same inputs -> same outputs is
           the whole parity contract; producing the history with 500-4000 CPU forwards would only burn GPU-box minutes.

Why teacher-forced and not "token stream identical": at 32 layers the synthetic gain-4 model amplifies fp32 rounding noise chaotically,
and the bf16 cache turns it into rounding-boundary flips. The REFERENCE ALGORITHM AGAINST ITSELF shows it: the C restatement built with
FMA contraction (oracle/_build/liboracle_port_fma.so = what `-O3 -march=x86-64-v3` does to the reference) leaves the strict build's
greedy stream at the 5th generated token on cfg4 (measured in the dev container: tokens 28505 vs 2143, margin 0.52); built with -Ofast
(liboracle_port_fast.so: serial sums re-associated into 8-lane partial sums, what every parallel implementation must do) its logits
move by up to 12 % of max|logit| on single positions of cfg4 with IDENTICAL inputs (position 517 of the test below). So (1) every position
is checked on IDENTICAL inputs: tokens are teacher-forced and so is the cache — after each forward the strict oracle's K/V rows of that
position replace the ones the engine (or the yardstick build) wrote, so nothing compounds from position to position; and (2) the
tolerance is CALIBRATED per position by the yardstick yard_i = max over the two builds of |logits_build_i - logits_strict_i| /
max|logits_strict_i| (the sensitivity is a property of the state at that position: both builds and the CUDA path peak at the same
positions): the CUDA path
must stay within max(5e-3, 8 * yard_i) of the strict oracle, i.e. no further from the reference than a legal re-association of the
reference's own arithmetic is (times 8: the envelope is two samples per position; measured yard_i ~ 1e-2 at 32 layers, and every
CUDA path measured between 0.5x and 4.2x of it, profiles/r02_full_config_parity.log). The K row
the engine writes for layer 0 (no depth amplification) must match the oracle's to a bf16 ulp. Every engine mode that serves the configuration is checked: the default megakernel(s), the word-based
megakernel (the kernel behind every N > 1 number) and the per-kernel CUDA-graph path. bf16 KV on both sides (orc_set_kv_bf16).
"""
import os

import numpy as np
import pytest
import torch

from conftest import oracle_shape
from oracle import loader
from simplellminference_b200.config import BF16, INT8, PRESETS
from simplellminference_b200.engine import Engine

pytestmark = pytest.mark.gpu

NT = os.cpu_count() or 1
MODES = {"mega": dict(mega=True), "mega_v2": dict(mega=True, mega_v2=True), "mega_v2_fuse": dict(mega=True, mega_v2=True, mega_fuse_down=True),
         "mega_ll": dict(mega=True, mega_ll=True), "fused_graph": {}}
ALL = ["mega", "mega_v2", "mega_v2_fuse", "mega_ll", "fused_graph"]
FLOOR, K_YARD = 5e-3, 8.0


def _need_ram(gib):
    try:
        avail = int(next(ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")).split()[1]) / 2**20
    except Exception:
        return
    if avail < gib:
        pytest.skip(f"host has {avail:.0f} GiB available, the fp32 oracle blob of this configuration needs {gib} GiB")


def _bf16_exact(a):
    return torch.from_numpy(a).bfloat16().float().numpy()


def _history(ms, n_pos, seed):
    """Per layer: K and V rows [n_pos][kv] ~ N(0,1), exactly representable in bf16 (so a bf16 and an fp32 cache hold the same values)."""
    rng = np.random.default_rng(seed)
    for l in range(ms.layers):
        yield l, _bf16_exact(rng.standard_normal((n_pos, ms.kv_hidden), dtype=np.float32)), _bf16_exact(rng.standard_normal((n_pos, ms.kv_hidden), dtype=np.float32))


def _inject_engine(eng, ms, l, k, v):
    n_pos = k.shape[0]
    for name, rows in (("key_cache", k), ("value_cache", v)):
        buf = eng.buffer(name)
        src = torch.from_numpy(rows).cuda().to(buf.dtype)
        if eng.lib.sllm_engine_kv_layout(eng.h) == 1:   # head-major [L][KVH][S][hd]
            buf.view(ms.layers, ms.kv_heads, ms.max_len, ms.head_dim)[l, :, :n_pos, :] = src.view(n_pos, ms.kv_heads, ms.head_dim).permute(1, 0, 2)
        else:                                           # the reference's [L][S][kv]
            buf.view(ms.layers, ms.max_len, ms.kv_hidden)[l, :n_pos, :] = src
    torch.cuda.synchronize()


def _rows_at(om, ms, pos):
    """The oracle's K and V cache rows of every layer at one position: ([L][kv], [L][kv])."""
    S, kv = ms.max_len, ms.kv_hidden
    k = np.stack([om.read(2, (l * S + pos) * kv, kv) for l in range(ms.layers)])
    v = np.stack([om.read(3, (l * S + pos) * kv, kv) for l in range(ms.layers)])
    return k, v


def _put_rows_oracle(om, ms, pos, k, v):
    S, kv = ms.max_len, ms.kv_hidden
    for l in range(ms.layers):
        om.write(2, (l * S + pos) * kv, k[l])
        om.write(3, (l * S + pos) * kv, v[l])


def _put_rows_engine(eng, ms, pos, k, v):
    for name, rows in (("key_cache", k), ("value_cache", v)):
        buf = eng.buffer(name)
        src = torch.from_numpy(rows).cuda().to(buf.dtype)
        if eng.lib.sllm_engine_kv_layout(eng.h) == 1:   # head-major [L][KVH][S][hd]
            buf.view(ms.layers, ms.kv_heads, ms.max_len, ms.head_dim)[:, :, pos, :] = src.view(ms.layers, ms.kv_heads, ms.head_dim)
        else:
            buf.view(ms.layers, ms.max_len, ms.kv_hidden)[:, pos, :] = src
    torch.cuda.synchronize()


def _oracle_runs(port, ms, wd, seed, group, n_hist, feed):
    """Strict oracle teacher-forced on `feed(i, strict_logits_so_far)`; then the FMA-contracted yardstick build (when the CPU can run
    it) on the same tokens AND the strict run's cache rows (so every position sees identical inputs on both sides).
    Returns (tokens fed, strict logits per position, strict K/V rows per position, yardstick per position or None)."""
    blob = port.fill_blob(oracle_shape(ms), seed, wd, group, threads=NT)
    S, kv = ms.max_len, ms.kv_hidden

    def model(lib):
        om = lib.model(oracle_shape(ms), blob, threads=NT, kv_bf16=True)
        if n_hist:
            for l, k, v in _history(ms, n_hist, seed + 1):
                om.write(2, l * S * kv, k)
                om.write(3, l * S * kv, v)
        return om

    om = model(port)
    toks, strict, rows, i = [], [], [], 0
    while True:
        t = feed(i, strict)
        if t is None:
            break
        toks.append(int(t))
        strict.append(om.forward(int(t), n_hist + i))
        rows.append(_rows_at(om, ms, n_hist + i))
        i += 1
    om.close()
    yard = None
    if loader.cpu_supports_v3():
        for so in (loader.PORT_FMA_SO, loader.PORT_FAST_SO):
            if not os.path.exists(so):
                continue
            om = model(loader.Port(so))
            y = []
            for i, t in enumerate(toks):
                if i:
                    _put_rows_oracle(om, ms, n_hist + i - 1, *rows[i - 1])
                lg = om.forward(t, n_hist + i)
                y.append(float(np.abs(lg - strict[i]).max()) / max(1.0, float(np.abs(strict[i]).max())))
            om.close()
            yard = y if yard is None else [max(a, b) for a, b in zip(yard, y)]
    del blob
    return toks, strict, rows, yard


def _check_engines(ms, wd, seed, group, n_hist, toks, strict, rows, yard, modes, what):
    report = {}
    for mode in modes:
        eng = Engine(ms, w_dtype=wd, kv_dtype=BF16, group=group, **MODES[mode]).load_synthetic(seed)
        if n_hist:
            for l, k, v in _history(ms, n_hist, seed + 1):
                _inject_engine(eng, ms, l, k, v)
        errs, kerrs = [], []
        for i, t in enumerate(toks):
            got, nxt = eng.forward(int(t), n_hist + i)
            want = strict[i]
            scale = max(1.0, float(np.abs(want).max()))
            err = float(np.abs(got - want).max())
            errs.append(err / scale)
            tol = max(FLOOR, K_YARD * yard[i]) if yard is not None else 5e-2
            assert err <= tol * scale, (what, mode, eng.mode, "position", n_hist + i, "rel err", err / scale, "tol", tol, "yardstick", yard[i] if yard else None)
            assert nxt == int(np.argmax(got))
            srt = np.partition(want, -2)[-2:]
            if srt[1] - srt[0] > 2.5 * err:
                assert nxt == int(np.argmax(want)), (what, mode, "arg-max at position", n_hist + i, err, float(srt[1] - srt[0]))
            # the K row the engine wrote for layer 0 at this position sees no depth amplification: one bf16 ulp at most
            k0 = eng.kv_row("k", 0, n_hist + i).float().cpu().numpy()
            kerrs.append(float(np.abs(k0 - rows[i][0][0]).max()) / max(1e-6, float(np.abs(rows[i][0][0]).max())))
            assert kerrs[-1] <= 1e-2, (what, mode, "layer-0 K row", n_hist + i, kerrs[-1])
            _put_rows_engine(eng, ms, n_hist + i, *rows[i])   # teacher-force the cache too: the next position sees the oracle's rows
        report[f"{mode}:{eng.mode}"] = dict(max_rel_err=f"{max(errs):.2e}", per_position=[f"{e:.1e}" for e in errs], layer0_k_row=f"{max(kerrs):.1e}")
        eng.close()
    print(f"\n[{what}] yardstick (reference vs its FMA-contracted build, identical inputs), per position:", None if yard is None else [f"{y:.1e}" for y in yard])
    for k, v in report.items():
        print(f"[{what}] {k}: {v}")
    return report


def _late(port, ms, wd, n_hist, tokens, modes, seed, what, group=64):
    toks, strict, rows, yard = _oracle_runs(port, ms, wd, seed, group, n_hist, lambda i, out: tokens[i] if i < len(tokens) else None)
    return _check_engines(ms, wd, seed, group, n_hist, toks, strict, rows, yard, modes, what)


def _head(port, ms, wd, prompt, n_total, modes, seed, what, group=64):
    def feed(i, out):   # predict's loop: prompt tokens verbatim, then the strict oracle's own arg-max
        if i >= n_total - 1:
            return None
        return prompt[i] if i < len(prompt) else int(np.argmax(out[i - 1]))
    toks, strict, rows, yard = _oracle_runs(port, ms, wd, seed, group, 0, feed)
    return _check_engines(ms, wd, seed, group, 0, toks, strict, rows, yard, modes, what)


# ---- cfg4: Llama-2-7B-shaped, 32 layers, bf16 weights + bf16 KV (the configuration the metric is quoted on) ---------------------------
def test_cfg4_llama2_7b_full_depth_head(port):
    _need_ram(34)
    _head(port, PRESETS["llama2-7b"], BF16, [1, 2, 3], 13, ALL, 1234, "cfg4 head")


def test_cfg4_llama2_7b_full_depth_positions_512_520(port):
    """Positions 512-520 after a 512-position history: the bench's position range (split-KV over > 512 positions, 32 KV heads)."""
    _need_ram(34)
    _late(port, PRESETS["llama2-7b"], BF16, 512, [11, 2222, 13000, 31999, 5, 777, 20481, 9, 1500], ALL, 1234, "cfg4 pos 512-520")


# ---- cfg5: Llama-3-8B-shaped, 32 layers, GQA-4, 128K vocabulary, position > 4096 --------------------------------------------------------
def test_cfg5_llama3_8b_full_depth_past_4096(port):
    _need_ram(38)
    _late(port, PRESETS["llama3-8b"], BF16, 4200, [7, 100000, 64000, 128255], ALL, 77, "cfg5 pos 4200-4203")


# ---- cfg3: TinyLlama-1.1B-shaped, 22 layers, int8 group-64 weights (and bf16), bf16 KV, near position 2000 ------------------------------
@pytest.mark.parametrize("wd", [INT8, BF16])
def test_cfg3_tinyllama_full_depth_near_2000(port, wd):
    modes = ["mega", "fused_graph"] if wd == INT8 else ALL
    _late(port, PRESETS["tinyllama-1.1b"], wd, 2000, [3, 31000, 15000, 42, 8191, 2], modes, 5, f"cfg3 wd={wd} pos 2000-2005")


def test_cfg3_tinyllama_full_depth_head_int8(port):
    _head(port, PRESETS["tinyllama-1.1b"], INT8, [1, 5, 9], 20, ["mega", "fused_graph"], 5, "cfg3 int8 head")
