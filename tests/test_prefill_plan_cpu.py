"""Host-side planning of the prefill GEMM (sllm_prefill_gemm_plan): pure arithmetic, no GPU needed. The plan decides the cluster
shape (one SM / an SM pair per tile), the N extent of a tile and the K split by a waves x operand-ingest cost model
(simplellminference_b200/csrc/prefill_gemm.cu pf_plan)."""
import ctypes as C

import pytest

from simplellminference_b200 import _lib


def plan(T, N, K, residual):
    lib = _lib.load()
    a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(lib.sllm_prefill_gemm_plan(T, N, K, int(residual), C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return a.value, b.value, c.value, d.value


@pytest.mark.parametrize("T", [1, 77, 128, 129, 512, 1000, 2048])
@pytest.mark.parametrize("N,K", [(128, 64), (1000, 1408), (4096, 4096), (12288, 4096), (22016, 4096), (4096, 11008), (32000, 4096)])
@pytest.mark.parametrize("residual", [False, True])
def test_plan_invariants(T, N, K, residual):
    two_sm, bn, ksplit, units = plan(T, N, K, residual)
    assert two_sm == int(T > 128)                                  # more than one 128-row block -> SM pairs (cta_group::2)
    assert bn % 32 == 0 and 64 <= bn <= 256                         # one tcgen05.mma N per tile, whole 32-column epilogue steps
    assert 1 <= ksplit <= 4 and (residual or ksplit == 1)           # only the residual epilogue can add partial sums
    bm = 256 if two_sm else 128
    assert units == -(-T // bm) * -(-N // bn) * ksplit
    if ksplit > 1:
        assert -(-K // 64) // ksplit >= 8                           # at least 8 k-blocks per work unit


def test_plan_llama2_7b_prompt_of_512():
    """The shapes of the headline prefill: 74 SM pairs."""
    assert plan(512, 12288, 4096, False)[1:3] == (192, 1)           # qkv: 2 x 64 tiles of 256 x 192 = 2 waves, 86 % full
    two_sm, bn, ksplit, units = plan(512, 4096, 4096, True)         # wo: one wave of 256 x 128 tiles left 10 pairs idle and 1.5 MB per CTA
    assert ksplit >= 2 and units <= 74
    two_sm, bn, ksplit, units = plan(512, 4096, 11008, True)        # down
    assert ksplit >= 2 and units <= 74
    assert plan(512, 4096, 4096, False)[2] == 1                     # the same GEMM under tensor parallelism (partial sums are stored)


def test_plan_rejects_bad_shapes():
    lib = _lib.load()
    assert lib.sllm_prefill_gemm_plan(0, 128, 64, 0, None, None, None, None) != 0
    assert lib.sllm_prefill_gemm_plan(128, 128, 60, 0, None, None, None, None) != 0     # K must be a multiple of 8 (16-byte TMA strides)
