"""GPU parity tests of the EXPERIMENTAL decode megakernel variant with the down projection fused into the gate_up phase
(SLLM_ENGINE_MEGA_FUSE_DOWN; csrc/megakernel.cu FUSE = true, megakernel.cuh "PH_DOWN_T"): a K-split of Wdown over the CTAs whose
partial output vectors are added into the residual stream with red.global.add.v4.f32 — four grid barriers per layer instead of
five. The summation order of the down projection is not fixed, so runs are not bit-reproducible; the contract is the decode
contract of tests/test_engine_gpu.py: token stream IDENTICAL to the reference / the oracle, logits within 3e-4 * max|logit|
(fp32 cache) or 5e-3 (bf16 cache).

(The file sorts after every other GPU test on purpose: a fault of this never-run kernel must not poison the CUDA context of verified tests.)

First ran (and passed) at the end of round 1 on the driver's B200; plain tests since then."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import oracle_shape
from simplellminference_b200.config import BF16, F32, PRESETS, ModelShape
from simplellminference_b200.engine import Engine

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


@pytest.mark.parametrize("d,inter,dt", [(128, 384, "f32"), (256, 704, "bf16"), (4096, 11008, "bf16"), (2048, 5632, "f32")])
def test_transposed_down_layout(d, inter, dt):
    """The repack kernel alone against the index formula of include/sllm_b200.h (and of tests/test_mega_fuse_layout_cpu.py): bit copy."""
    import ctypes as C
    import torch
    from simplellminference_b200 import _lib
    lib = _lib.load()
    tdt, E, wd = (torch.float32, 4, F32) if dt == "f32" else (torch.bfloat16, 8, BF16)
    W = torch.randn(d, inter, device="cuda").to(tdt).contiguous()
    out = torch.empty(d * inter, device="cuda", dtype=tdt)
    _lib.check(lib.sllm_mega_repack_down_t(W.data_ptr(), out.data_ptr(), d, inter, wd, None))
    torch.cuda.synchronize()
    i = torch.arange(d * inter, device="cuda")
    ntr = inter // 4
    e, lane, jj = i % E, (i // E) % 32, (i // (E * 32)) % 4
    g, ks = (i // (E * 32 * 4)) % ntr, i // (E * 32 * 4 * ntr)
    want = W.reshape(-1)[((ks * 32 + lane) * E + e) * inter + (g * 4 + jj)]
    assert torch.equal(out.view(torch.int16 if dt == "bf16" else torch.int32), want.view(torch.int16 if dt == "bf16" else torch.int32))
    assert lib.sllm_mega_repack_down_t(W.data_ptr(), out.data_ptr(), 768, 2048, F32, None) != 0     # six stripes: not a power of two


def test_golden_stream_of_the_reference():
    """tiny_gqa fp32 (hidden 128 = one 512-byte stripe, 16 row groups): the token stream and logits recorded from the UNMODIFIED reference."""
    golden = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "models_ref.npz")))
    prompt, n_total, wd = mg.MODEL_RUNS["tiny_gqa"]
    eng = Engine(PRESETS["tiny_gqa"], w_dtype=wd, kv_dtype=F32, mega=True, mega_fuse_down=True).load_synthetic(mg.SEED)
    assert eng.mode == "megakernel(fused-down)", eng.mode
    toks = eng.greedy(prompt, n_total)
    want, want_l = golden["tiny_gqa/tokens"], golden["tiny_gqa/last_logits"]
    assert np.array_equal(toks, want), (np.flatnonzero(toks != want)[:5], toks[:8], want[:8])
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
    assert err <= 3e-4 * max(1.0, float(np.abs(want_l).max())), err
    x_last, want_x = eng.buffer("emb_output").cpu().numpy(), golden["tiny_gqa/x_last"]
    assert float(np.abs(x_last - want_x).max()) <= 3e-4 * max(1.0, float(np.abs(want_x).max()))
    eng.close()


@pytest.mark.parametrize("d,heads,kvh,inter,wd", [(256, 4, 2, 704, BF16), (512, 8, 8, 1408, BF16), (1024, 16, 4, 2824, BF16), (256, 4, 4, 516, F32)])
def test_stripe_counts_against_the_oracle(port, d, heads, kvh, inter, wd):
    """1 / 2 / 4 stripes of 512 bytes per output row (16 / 8 / 4 warps share a stripe and split the tile rows), an intermediate size
    that leaves CTAs with different unit counts, bf16 and fp32 weights; fp32 cache so that identity does not hinge on bf16 rounding."""
    ms = ModelShape(1000, d // heads, d, kvh * (d // heads), inter, 40, 2, heads, kvh)
    blob = port.fill_blob(oracle_shape(ms), 17, wd, 64)
    want, want_l = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1).greedy([1, 2, 3], 38)
    eng = Engine(ms, w_dtype=wd, kv_dtype=F32, mega=True, mega_fuse_down=True).load_blob(blob)
    assert eng.mode == "megakernel(fused-down)", eng.mode
    got = eng.greedy([1, 2, 3], 38)
    assert np.array_equal(got, want), (int(np.flatnonzero(got != want)[0]), got[:8], want[:8])
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
    assert err <= 3e-4 * max(1.0, float(np.abs(want_l).max())), err
    plain = Engine(ms, w_dtype=wd, kv_dtype=F32, mega=True).load_blob(blob)          # the verified kernel on the same weights
    assert plain.mode == "megakernel" and np.array_equal(plain.greedy([1, 2, 3], 38), want)
    plain.close(); eng.close()


def test_full_width_llama2_7b_two_layers(port):
    """The widths of the bench (hidden 4096 = 16 stripes, one per warp; 11008 inputs = 2752 tile rows of four over 148 CTAs) with two
    layers against the oracle (bf16 cache rows on both sides), then the batched prefill on the same engine: it keeps reading the
    standard tiled down matrix."""
    import dataclasses
    ms = dataclasses.replace(PRESETS["llama2-7b"], layers=2, max_len=64)
    blob = port.fill_blob(oracle_shape(ms), 9, BF16, 64, threads=os.cpu_count() or 1)
    want, want_l = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1, kv_bf16=True).greedy([1, 2, 3], 14)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, mega=True, mega_fuse_down=True).load_synthetic(9)
    assert eng.mode == "megakernel(fused-down)", eng.mode
    got = eng.greedy([1, 2, 3], 14)
    assert np.array_equal(got, want), (got, want)
    err = float(np.abs(eng.buffer("model_pred").cpu().numpy() - want_l).max())
    assert err <= 5e-3 * max(1.0, float(np.abs(want_l).max())), err
    assert eng.prefill_supported
    ids = np.concatenate([[1, 2, 3], want[:9]]).astype(np.int32)
    eng.prefill(ids)
    eng.enqueue_steps(2)
    assert np.isfinite(eng.buffer("model_pred").cpu().numpy()).all()
    eng.close()


def test_shapes_it_does_not_take_fall_back_visibly():
    """hidden 768 fp32 = six stripes (not a power of two), int8 weights: the flag is ignored and the mode string says so."""
    from simplellminference_b200.config import INT8
    for ms, wd in ((PRESETS["stories110M"], F32), (PRESETS["tiny_gqa"], INT8)):
        eng = Engine(ms, w_dtype=wd, kv_dtype=F32, mega=True, mega_fuse_down=True)
        assert eng.mode == "megakernel", eng.mode
        eng.close()
