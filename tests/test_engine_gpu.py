"""GPU parity tests, model level: the engine (fused CUDA-graph path, and the op-by-op path) against the golden
token streams/logits produced by the unmodified reference, and against the CPU oracle on fresh seeds.

Decode tolerance: token sequence IDENTICAL; max|dlogit| <= 3e-4 * max(1, max|logit|) at the end of a greedy run
of up to 256 tokens, 1e-4 for a single forward from oracle-identical history; and the oracle's top1-top2 margin
must dwarf the observed logit error so identity is meaningful. Why 3e-4 and not SURVEY.md's 1e-4: everything is
fp32 on both sides and only the summation order differs, but the synthetic projections have gain 4
(std 4/sqrt(fan_in)), which amplifies rounding noise layer by layer. The yardstick is the reference against
ITSELF: its -O2 -ffp-contract=off and -O3 -march=x86-64-v3 (FMA-contracted) builds differ by 8.1e-5 * max|logit|
in the final logits of the cfg2 run below (measured in the dev container, see DESIGN.md "Tolerances").
Residual stream / KV rows are held to the same relative bound against their own max magnitude."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import oracle_shape
from simplellminference_b200.config import F32, BF16, INT8, PRESETS, ModelShape
from simplellminference_b200.engine import Engine

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

PRESET_OF = {"cfg1_stories15M": "stories15M", "cfg2_stories110M": "stories110M", "cfg2_stories110M_256": "stories110M", "tiny_gqa": "tiny_gqa",
             "tiny_gqa_bf16w": "tiny_gqa", "tiny_gqa_int8w": "tiny_gqa", "tiny_mha_hd48": "tiny_mha_hd48"}


def logit_tol(want):
    return 3e-4 * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("name", list(mg.MODEL_RUNS))
@pytest.mark.parametrize("mode", ["mega", "mega_v2", "mega_v2_fuse", "mega_ll", "fused_graph", "fused_graph_pdl", "fused_nograph", "unfused"])
def test_golden_models(golden_models, name, mode):
    """Token streams and final logits recorded from the reference itself (tests/golden/models_ref.npz)."""
    prompt, n_total, wd = mg.MODEL_RUNS[name]
    ms = PRESETS[PRESET_OF[name]]
    kw = dict(mega=dict(mega=True), mega_v2=dict(mega=True, mega_v2=True), mega_v2_fuse=dict(mega=True, mega_v2=True, mega_fuse_down=True),
              mega_ll=dict(mega=True, mega_ll=True), fused_graph={}, fused_graph_pdl=dict(pdl=True), fused_nograph=dict(graph=False), unfused=dict(fused=False))[mode]
    eng = Engine(ms, w_dtype=wd, kv_dtype=F32, group=64, **kw).load_synthetic(mg.SEED)
    if mode.startswith("mega"):
        rb = ms.hidden * (4 if wd == F32 else 2)   # mega_fuse_down_ok: a row of outputs = 2^k stripes of 512 bytes
        fuse_ok = wd != INT8 and rb % 512 == 0 and (rb // 512) & (rb // 512 - 1) == 0 and rb // 512 <= 16 and ms.inter % 4 == 0
        want_mode = {"mega": "megakernel", "mega_ll": "megakernel(ll)", "mega_v2": "megakernel(v2)",
                     "mega_v2_fuse": "megakernel(v2,fused-down)" if fuse_ok else "megakernel(v2)"}[mode]
        if wd == INT8:   # int8 group-64 tiles: the grid-barrier kernel and the word-based one stream them, megakernel2 falls back (visibly)
            want_mode = "megakernel(ll)" if mode == "mega_ll" else "megakernel"
        assert eng.mode == want_mode, eng.mode
    want = golden_models[name + "/tokens"]
    want_l = golden_models[name + "/last_logits"]
    if name.endswith("_256"):
        # BASELINE.json configs[1] at its full 256 tokens. Over such a run the fp32 rounding noise of ANY re-ordered sum compounds through
        # the cache (every CUDA path alike: 3e-3 of max|logit| by position 120, 1e-2 by position 170 against the oracle on identical
        # tokens, tools/v2_locate.py), while the reference's own stream has top-1/top-2 margins down to 0.02 (position 159): identity
        # beyond ~140 tokens is luck. Held to: the first 128 tokens identical in free-running greedy decode, and — teacher-forced on
        # the reference's tokens — the echo of all 255 and the final logits within 5e-2 of max|logit|.
        toks = eng.greedy(prompt, n_total)
        assert np.array_equal(toks[:128], want[:128]), (np.flatnonzero(toks[:128] != want[:128])[:5],)
        print(f"\n{name} [{mode}]: free-running greedy stream identical to the reference's for {int(np.argmax(toks != want)) if (toks != want).any() else toks.size} of {want.size} tokens")
        forced = eng.greedy(np.concatenate([prompt, want[:-1]]).astype(np.int32), n_total)
        assert np.array_equal(forced[:-1], want[:-1])
        err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
        assert err <= 5e-2 * float(np.abs(want_l).max()), err
        eng.close()
        return
    toks = eng.greedy(prompt, n_total)
    assert np.array_equal(toks, want), (np.flatnonzero(toks != want)[:5], toks[:8], want[:8])
    logits = eng.buffer("model_pred").cpu().numpy()
    err = float(np.abs(logits - want_l).max())
    assert err <= logit_tol(want_l), err
    srt = np.sort(want_l)
    assert srt[-1] - srt[-2] >= 20 * err or err < 1e-5
    k_last = eng.kv_row("k", 0, n_total - 2).float().cpu().numpy()
    want_k = golden_models[name + "/k_l0_last"]
    assert float(np.abs(k_last - want_k).max()) <= logit_tol(want_k)
    x_last, want_x = eng.buffer("emb_output").cpu().numpy(), golden_models[name + "/x_last"]
    assert float(np.abs(x_last - want_x).max()) <= logit_tol(want_x)
    eng.close()


def test_blob_loader_equals_synthetic(port):
    """Loading the reference-format fp32 blob from the host == generating on the device (all three dtypes)."""
    ms = PRESETS["tiny_gqa"]
    for wd in (F32, BF16, INT8):
        blob = port.fill_blob(oracle_shape(ms), 1234, F32)       # raw fp32: the engine converts
        a = Engine(ms, w_dtype=wd, kv_dtype=F32).load_blob(blob)
        b = Engine(ms, w_dtype=wd, kv_dtype=F32).load_synthetic(1234)
        for bid in (100, 101, 102, 103, 104, 105):
            ta, tb = a.buffer(bid), b.buffer(bid)
            assert torch.equal(ta.view(torch.uint8) if ta.dtype == torch.bfloat16 else ta, tb.view(torch.uint8) if tb.dtype == torch.bfloat16 else tb), (wd, bid)
        assert np.array_equal(a.greedy([1, 2, 3], 20), b.greedy([1, 2, 3], 20))
        a.close(); b.close()


@pytest.mark.parametrize("mega", [True, False, "v2"])
def test_forward_api_matches_oracle_per_position(port, mega):
    """LlamaModel::forward semantics: explicit (token, pos), logits back on the host, every position checked."""
    ms = PRESETS["tiny_mha_hd48"]
    blob = port.fill_blob(oracle_shape(ms), 42)
    om = port.model(oracle_shape(ms), blob)
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32, mega=bool(mega), mega_v2=(mega == "v2")).load_blob(blob)
    assert mega != "v2" or eng.mode == "megakernel(v2)", eng.mode
    tok = 5
    for pos in range(ms.max_len):
        want = om.forward(tok, pos)
        got, nxt = eng.forward(tok, pos)
        assert float(np.abs(got - want).max()) <= 1e-4 * max(1.0, float(np.abs(want).max()))
        assert nxt == int(np.argmax(want))
        tok = nxt
    eng.close()


@pytest.mark.parametrize("mega", [True, False, "v2", "v2fuse"])
@pytest.mark.parametrize("wd,kvd", [(BF16, BF16), (INT8, BF16), (F32, BF16)])
def test_bf16_kv_cache_variant(port, wd, kvd, mega):
    """bf16 KV cache has no reference implementation: its definition is the oracle with cache rows rounded to
    bf16 (orc_set_kv_bf16). Tokens identical, logits within the decode tolerance."""
    ms = PRESETS["tiny_gqa"]
    blob = port.fill_blob(oracle_shape(ms), 7, wd, 64)
    om = port.model(oracle_shape(ms), blob, kv_bf16=True)
    want, want_l = om.greedy([1, 9], 46)
    eng = Engine(ms, w_dtype=wd, kv_dtype=kvd, mega=bool(mega), mega_v2=str(mega).startswith("v2"), mega_fuse_down=(mega == "v2fuse")).load_synthetic(7)
    got = eng.greedy([1, 9], 46)
    logits = eng.buffer("model_pred").cpu().numpy()
    err = float(np.abs(logits - want_l).max())
    srt = np.sort(want_l)
    assert np.array_equal(got, want), (err, srt[-1] - srt[-2])
    assert err <= 5e-3 * max(1.0, float(np.abs(want_l).max())), err   # a cache value on a bf16 rounding boundary may flip
    eng.close()


def test_full_context_and_state_api(port):
    """Run to the last position of the parity domain (pos <= S - H/KVH); device-resident stepping API."""
    ms = PRESETS["tiny_gqa"]
    blob = port.fill_blob(oracle_shape(ms), 11)
    want, _ = port.model(oracle_shape(ms), blob).greedy([3], 47)
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32).load_synthetic(11)
    eng.set_state(3, 0)
    eng.enqueue_steps(20)
    eng.enqueue_steps(26)
    assert np.array_equal(eng.read_tokens(46), want)
    with pytest.raises(Exception):
        eng.enqueue_steps(10)     # would overrun max_len
    assert eng.step_launches == 2 + 5 * ms.layers
    eng.close()


@pytest.mark.parametrize("mega", [True, False, "v2", "v2fuse"])
def test_medium_shape_tokens(port, mega):
    """A GQA shape with the production head_dim (128) and a long-ish context, bf16 weights, fp32 KV."""
    ms = ModelShape(4096, 128, 1024, 256, 2816, 160, 4, 8, 2)
    blob = port.fill_blob(oracle_shape(ms), 3, BF16)
    want, want_l = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1).greedy(list(range(1, 33)), 150)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=F32, mega=bool(mega), mega_v2=str(mega).startswith("v2"), mega_fuse_down=(mega == "v2fuse")).load_synthetic(3)
    # fp32 cache rows of 128-wide heads (the parity-mode cache on the production shapes): in the megakernels since round 2 (K/V stages of 32 positions)
    assert eng.mode == {True: "megakernel", False: "fused+graph", "v2": "megakernel(v2)", "v2fuse": "megakernel(v2,fused-down)"}[mega], eng.mode
    got = eng.greedy(list(range(1, 33)), 150)
    assert np.array_equal(got, want)
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
    assert err <= logit_tol(want_l), err
    eng.close()


@pytest.mark.parametrize("kvd", [F32, BF16])
@pytest.mark.parametrize("mega_ll", [False, True, "v2", "v2fuse"])
def test_gqa8_hd64_megakernel(port, kvd, mega_ll):
    """TinyLlama's head geometry (8 query heads per KV head, head_dim 64): the cross-stripe attention buffer spans both K/V
    stages. Must run in megakernel mode (it used to fall back) and match the oracle."""
    ms = ModelShape(2048, 64, 512, 64, 1408, 200, 3, 8, 1)
    blob = port.fill_blob(oracle_shape(ms), 5, BF16)
    want, want_l = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1, kv_bf16=(kvd == BF16)).greedy([1, 5, 9], 190)
    v2 = str(mega_ll).startswith("v2")
    eng = Engine(ms, w_dtype=BF16, kv_dtype=kvd, mega=True, mega_ll=(mega_ll is True), mega_v2=v2, mega_fuse_down=(mega_ll == "v2fuse")).load_synthetic(5)
    assert eng.mode == {False: "megakernel", True: "megakernel(ll)", "v2": "megakernel(v2)", "v2fuse": "megakernel(v2,fused-down)"}[mega_ll], eng.mode
    got = eng.greedy([1, 5, 9], 190)
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
    if kvd == F32:
        assert np.array_equal(got, want), int(np.flatnonzero(got != want)[0])
        assert err <= logit_tol(want_l), err
    else:   # bf16 cache: a value on a rounding boundary may flip and move a low-margin token late in a long run
        same = int(np.flatnonzero(got != want)[0]) if not np.array_equal(got, want) else len(want)
        assert same >= 60, same
    eng.close()


def test_int8_megakernel_shapes(port):
    """int8 group-64 weights in the persistent megakernel on shapes that exercise the tile geometry: K slices whose chunk count
    had to be rounded up to a whole quantisation group (inter = 1408 -> 88 chunks over 2 slices), 2- and 4-row tiles, GQA."""
    for ms, seed in ((ModelShape(2048, 64, 512, 128, 1408, 96, 3, 8, 2), 3), (ModelShape(1024, 64, 1024, 256, 2816, 80, 2, 16, 4), 4)):
        blob = port.fill_blob(oracle_shape(ms), seed, INT8, 64)
        want, want_l = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1).greedy([1, 2, 3, 4], 70)
        eng = Engine(ms, w_dtype=INT8, kv_dtype=F32, group=64, mega=True).load_synthetic(seed)
        assert eng.mode == "megakernel", eng.mode
        got = eng.greedy([1, 2, 3, 4], 70)
        assert np.array_equal(got, want), (int(np.flatnonzero(got != want)[0]), got[:8], want[:8])
        err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
        assert err <= logit_tol(want_l), err
        ref = Engine(ms, w_dtype=INT8, kv_dtype=F32, group=64, mega=False).load_synthetic(seed)   # per-kernel int8 path: same tokens
        assert np.array_equal(ref.greedy([1, 2, 3, 4], 70), want)
        ll = Engine(ms, w_dtype=INT8, kv_dtype=F32, group=64, mega=True, mega_ll=True).load_synthetic(seed)   # the word-based kernel (the tensor-parallel one)
        assert ll.mode == "megakernel(ll)", ll.mode
        got = ll.greedy([1, 2, 3, 4], 70)
        assert np.array_equal(got, want), (int(np.flatnonzero(got != want)[0]), got[:8], want[:8])
        assert float(np.abs(ll.buffer("model_pred").cpu().numpy() - want_l).max()) <= logit_tol(want_l)
        eng.close(); ref.close(); ll.close()


@pytest.mark.parametrize("wd,mega", [(BF16, True), (INT8, True), (BF16, False), (BF16, "v2"), (BF16, "v2fuse"), (INT8, "ll")])
def test_full_width_llama2_7b_two_layers(port, wd, mega):
    """The Llama-2-7B WIDTHS (d 4096, inter 11008, vocab 32000, 32 heads of 128) with 2 layers: the tile geometry the bench runs on
    (K = 11008 -> 16 K slices of 86 chunks, 2-row tiles; int8: 88 chunks) checked against the oracle, which the small shapes cannot
    do. bf16 KV cache (fp32 K/V tiles of 128-wide heads do not fit the megakernel), so the oracle rounds its cache rows too."""
    import dataclasses
    ms = dataclasses.replace(PRESETS["llama2-7b"], layers=2, max_len=64)
    # seed 10: the oracle's top-1/top-2 margins along this stream are >= 5.8 (on logits of ~280). Seed 9, used until round 2, has a
    # 0.0018 margin at the first generated token: the grid-barrier kernel itself lands on either side of it depending on max_len
    # (the number of attention splits), so a token-identity assertion there tested luck, not parity
    blob = port.fill_blob(oracle_shape(ms), 10, wd, 64, threads=os.cpu_count() or 1)
    om = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1, kv_bf16=True)
    want, want_l = om.greedy([1, 2, 3], 14)
    eng = Engine(ms, w_dtype=wd, kv_dtype=BF16, group=64, mega=bool(mega), mega_v2=str(mega).startswith("v2"), mega_fuse_down=(mega == "v2fuse"),
                 mega_ll=(mega == "ll")).load_synthetic(10)
    assert eng.mode == {True: "megakernel", False: "fused+graph", "v2": "megakernel(v2)", "v2fuse": "megakernel(v2,fused-down)", "ll": "megakernel(ll)"}[mega], eng.mode
    got = eng.greedy([1, 2, 3], 14)
    assert np.array_equal(got, want), (got, want)
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
    assert err <= 5e-3 * max(1.0, float(np.abs(want_l).max())), err
    if wd == BF16 and mega is True:   # the same prompt as ONE batched tensor-core pass: K = 11008 with its ragged k-block tail per K slice
        ids = np.concatenate([[1, 2, 3], want[:9]]).astype(np.int32)
        om2 = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1, kv_bf16=True)
        for p in range(ids.size - 1):
            om2.step(int(ids[p]), p)
        pl = om2.forward(int(ids[-1]), ids.size - 1)
        eng.prefill(ids)
        torch.cuda.synchronize()
        perr = float(np.abs(eng.buffer("model_pred").cpu().numpy() - pl).max())
        assert perr <= 6e-2 * max(1.0, float(np.abs(pl).max())), perr     # bf16 operands, 2 layers of the gain-4 model
        k1 = eng.kv_row("k", 1, 5).float().cpu().numpy()
        wk1 = om2.read(2, (1 * ms.max_len + 5) * ms.kv_hidden, ms.kv_hidden)
        assert float(np.abs(k1 - wk1).max()) <= 3e-2 * float(np.abs(wk1).max())
    eng.close()


@pytest.mark.parametrize("mega_ll", [False, True, "v2", "v2fuse"])
def test_full_width_llama3_8b_one_layer(port, mega_ll):
    """Llama-3-8B widths (GQA 4 query heads per KV head, inter 14336 -> 112-chunk K slices, 128 256-row tied classifier) with one
    layer, both megakernels, against the oracle (bf16 KV on both sides)."""
    import dataclasses
    ms = dataclasses.replace(PRESETS["llama3-8b"], layers=1, max_len=48)
    blob = port.fill_blob(oracle_shape(ms), 13, BF16, 64, threads=os.cpu_count() or 1)
    want, want_l = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1, kv_bf16=True).greedy([1, 2, 3, 4, 5], 14)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, mega=True, mega_ll=(mega_ll is True), mega_v2=str(mega_ll).startswith("v2"), mega_fuse_down=(mega_ll == "v2fuse")).load_synthetic(13)
    assert eng.mode == {False: "megakernel", True: "megakernel(ll)", "v2": "megakernel(v2)", "v2fuse": "megakernel(v2,fused-down)"}[mega_ll], eng.mode
    got = eng.greedy([1, 2, 3, 4, 5], 14)
    assert np.array_equal(got, want), (got, want)
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
    assert err <= 5e-3 * max(1.0, float(np.abs(want_l).max())), err
    eng.close()


def test_predict_driver(port):
    """simplellminference_b200.predict.predict_ids (the loop of LlamaModel::predict, model.cpp:148-185) around a real engine:
    identical to the oracle's greedy stream token by token; with the batched prefill the prompt echo is exact and the stream is the
    oracle's on the gain-1 model; EOS stops generation (additive — the reference never stops)."""
    from simplellminference_b200.predict import predict_ids
    ms = PRESETS["tiny_gqa"]
    blob = port.fill_blob(oracle_shape(ms), 1234)
    want, _ = port.model(oracle_shape(ms), blob).greedy([1, 7, 300], 41)
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32, mega=True).load_blob(blob)
    got = predict_ids(eng, [1, 7, 300], 40, chunk=7)
    assert np.array_equal(got, want)
    eos = int(want[20])
    first = int(np.flatnonzero(want == eos)[0])
    stopped = predict_ids(eng, [1, 7, 300], 40, eos_id=eos, chunk=7)
    assert np.array_equal(stopped, want[:first + 1])
    eng.close()

    ms2 = ModelShape(2048, 64, 512, 128, 1408, 320, 2, 8, 2)
    sh2 = oracle_shape(ms2)
    blob2 = port.fill_blob(sh2, 21, BF16)
    for seg in range(2, 9):
        off, cnt = port.segment(sh2, seg)[:2]
        blob2[off:off + cnt] *= 0.25
    prompt = np.random.default_rng(2).integers(1, ms2.vocab, size=100, dtype=np.int32)
    want2, _ = port.model(sh2, blob2, threads=os.cpu_count() or 1, kv_bf16=True).greedy(prompt, 121)
    eng2 = Engine(ms2, w_dtype=BF16, kv_dtype=BF16, mega=True).load_blob(blob2)
    assert eng2.prefill_supported
    got2 = predict_ids(eng2, prompt, 120, batched_prefill=True)   # opt-in tensor-core prompt pass (bf16 operands): gain-1 model, see DESIGN.md
    assert np.array_equal(got2[:99], prompt[1:])
    assert np.array_equal(got2, want2), int(np.flatnonzero(got2 != want2)[0])
    eng2.close()


@pytest.mark.parametrize("v2", [False, True])
def test_calibrated_partition_keeps_results(port, v2):
    """sllm_engine_calibrate re-cuts every phase's tile rows among the CTAs by their measured streaming rate: WHICH CTA computes a row
    changes, how it is computed does not — the deterministic kernel must reproduce its logits bit for bit, the v2 kernel (red.add
    order not fixed) its tokens and the decode tolerance, and both the oracle's stream."""
    ms = ModelShape(2048, 64, 512, 128, 1408, 96, 3, 8, 2)
    blob = port.fill_blob(oracle_shape(ms), 21, BF16)
    want, want_l = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1).greedy([1, 2, 3, 4], 60)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=F32, mega=True, mega_v2=v2, mega_fuse_down=v2).load_synthetic(21)
    assert eng.mode == ("megakernel(v2,fused-down)" if v2 else "megakernel"), eng.mode
    before = eng.greedy([1, 2, 3, 4], 60)
    logits_before = eng.buffer("model_pred").cpu().numpy().copy()
    assert eng.calibration().size == 0
    eng.calibrate(2)
    tau = eng.calibration()
    assert tau.size > 0 and abs(float(tau.mean()) - 1.0) < 0.05 and float(tau.min()) >= 0.8 and float(tau.max()) <= 1.25, tau
    after = eng.greedy([1, 2, 3, 4], 60)
    logits_after = eng.buffer("model_pred").cpu().numpy()
    assert np.array_equal(before, want) and np.array_equal(after, want)
    if not v2:
        assert np.array_equal(logits_before, logits_after)
    assert float(np.abs(logits_after - want_l).max()) <= logit_tol(want_l)
    eng.close()


@pytest.mark.parametrize("mega", [True, "v2fuse", "ll"])
def test_fp32_kv_wide_heads_long_context(port, mega):
    """fp32 cache + 128-wide heads (512-byte rows: K/V stages of 32 positions) far enough into the context that a split spans several
    stages (the tile loop, not the single-pass path): 8 kv heads -> 18 splits of ~100 positions at position 1800. Teacher-forced on a
    random prompt, then the logits of the last prompt position and 12 greedy tokens against the oracle."""
    ms = ModelShape(2048, 128, 1024, 1024, 1408, 1900, 2, 8, 8)
    blob = port.fill_blob(oracle_shape(ms), 17, BF16)
    prompt = np.random.default_rng(4).integers(1, ms.vocab, size=1800, dtype=np.int32)
    om = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1)
    want, want_l = om.greedy(prompt, 1812)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=F32, mega=True, mega_v2=(mega == "v2fuse"), mega_fuse_down=(mega == "v2fuse"), mega_ll=(mega == "ll")).load_synthetic(17)
    assert eng.mode == {True: "megakernel", "v2fuse": "megakernel(v2,fused-down)", "ll": "megakernel(ll)"}[mega], eng.mode
    got = eng.greedy(prompt, 1812)
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max()) / max(1.0, float(np.abs(want_l).max()))
    same = int(np.argmax(got != want)) if (got != want).any() else got.size
    print(f"\nfp32 KV, hd 128, 1811 positions [{eng.mode}]: identical prefix {same}/{got.size}, final logits within {err:.1e} of max|logit|")
    assert same >= 1799 + 6, same          # the echo of the prompt and at least the first generated tokens
    assert err <= 5e-3, err
    eng.close()
