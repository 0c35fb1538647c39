"""Host-side plan of the persistent decode megakernel (sllm_mega_plan, sllm_mega_tile_geometry): pure arithmetic, no GPU
needed once the two device facts (SM count, opt-in shared memory per block) are passed in. The plan decides whether the
one-launch-per-token kernel can take a shape (otherwise the engine runs the per-kernel fused path), its shared-memory
budget and the KV splits of its attention phase; the geometry is the tiled weight layout whose tile is one TMA bulk copy
into one ring slot (simplellminference_b200/csrc/megakernel.cu mega_plan_for / mega_tile_geom)."""
import ctypes as C
import dataclasses

import pytest

from simplellminference_b200 import _lib
from simplellminference_b200.config import BF16, F32, INT8, PRESETS

B200_SMS, B200_SMEM = 148, 232448        # B200: 148 SMs, 227 KB opt-in shared memory per block
SLOT_BYTES, WARPS = 4096, 16             # csrc/mega_common.cuh kSlotBytes, kMegaWarps
E = {F32: 4, BF16: 8, INT8: 16}


def c_shape(ms):
    return _lib.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)


def plan(ms, wd, kvd, tp=1, group=64, sms=B200_SMS, smem=B200_SMEM, word_based=None):
    """word_based None: what the engine picks (the grid-barrier kernel on one GPU, the word-based one under tensor parallelism)"""
    lib = _lib.load()
    ok, grid, nsplit, nbytes = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    wb = int(tp > 1) if word_based is None else int(word_based)
    _lib.check(lib.sllm_mega_plan(C.byref(c_shape(ms)), wd, group, kvd, tp, wb, sms, smem, C.byref(ok), C.byref(grid), C.byref(nbytes), C.byref(nsplit)))
    why = lib.sllm_last_error().decode() if not ok.value else ""
    return bool(ok.value), grid.value, nbytes.value, nsplit.value, why


def geometry(rows, cols, kind, wd):
    lib = _lib.load()
    ks, sc, r, ntr, tb, mb = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    _lib.check(lib.sllm_mega_tile_geometry(rows, cols, kind, wd, C.byref(ks), C.byref(sc), C.byref(r), C.byref(ntr), C.byref(tb), C.byref(mb)))
    return ks.value, sc.value, r.value, ntr.value, tb.value, mb.value


@pytest.mark.parametrize("name,wd,kvd,tp", [
    ("stories110M", F32, F32, 1), ("tinyllama-1.1b", BF16, BF16, 1), ("tinyllama-1.1b", INT8, BF16, 1),
    ("llama2-7b", BF16, BF16, 1), ("llama2-7b", INT8, BF16, 1), ("llama2-7b", BF16, BF16, 2), ("llama2-7b", BF16, BF16, 4),
    ("llama2-7b", BF16, BF16, 8), ("llama3-8b", BF16, BF16, 1), ("llama3-8b", BF16, BF16, 8)])
def test_benchmark_configurations_run_as_one_launch_per_token(name, wd, kvd, tp):
    """Every BASELINE.json configuration that is benchmarked through the megakernel gets a plan on a B200: one CTA per SM, within
    the opt-in shared memory, and at most one 64-position tile per KV split at full context."""
    ms = PRESETS[name]
    ok, grid, nbytes, nsplit, why = plan(ms, wd, kvd, tp)
    assert ok, why
    assert grid == B200_SMS and 0 < nbytes <= B200_SMEM
    kvh_loc = ms.kv_heads // tp
    assert 1 <= nsplit <= 32 and nsplit <= max(1, B200_SMS // kvh_loc) and nsplit <= -(-ms.max_len // 64)
    if kvh_loc * 32 >= B200_SMS and ms.max_len >= 64 * (B200_SMS // kvh_loc):
        assert nsplit == B200_SMS // kvh_loc          # the attention items fill the grid once


def test_headline_configuration_numbers():
    """llama2-7b bf16/bf16 on one B200: 4 KV splits per head (128 attention items over 148 CTAs); the ring is 16 warps x 2 slots x 4 KB."""
    ok, grid, nbytes, nsplit, _ = plan(PRESETS["llama2-7b"], BF16, BF16)
    assert ok and nsplit == 4
    assert nbytes >= WARPS * 2 * SLOT_BYTES + 2 * 2 * 64 * 128 * 2      # the ring + two K and two V stages of 64 rows of 128 bf16


def test_shapes_the_megakernel_declines_say_why():
    ms = PRESETS["llama2-7b"]
    ok, *_, why = plan(ms, INT8, BF16, group=32)
    assert not ok and "group" in why
    ok, grid, nbytes, *_ = plan(ms, BF16, F32)         # fp32 cache rows of 128-wide heads (the parity-mode cache on the 7B / 8B shapes): taken since
    assert ok and nbytes <= B200_SMEM                  # round 2 — a K/V stage holds 32 positions instead of 64 when a row is wider than 256 bytes
    ok, *_, why = plan(ms, BF16, BF16, smem=48 * 1024)
    assert not ok and "shared memory" in why
    odd = dataclasses.replace(PRESETS["tiny_gqa"], hidden=132, heads=4, head_dim=33, kv_hidden=66)
    ok, *_, why = plan(odd, BF16, BF16)
    assert not ok and why
    lib = _lib.load()
    one = C.c_int32()
    assert lib.sllm_mega_plan(C.byref(c_shape(ms)), BF16, 64, BF16, 3, 1, B200_SMS, B200_SMEM, C.byref(one), None, None, None) != 0   # 32 heads over 3 ranks
    assert lib.sllm_mega_plan(C.byref(c_shape(ms)), BF16, 64, BF16, 2, 0, B200_SMS, B200_SMEM, C.byref(one), None, None, None) != 0   # TP needs the word-based kernel
    assert lib.sllm_mega_plan(None, BF16, 64, BF16, 1, 0, B200_SMS, B200_SMEM, C.byref(one), None, None, None) != 0


def test_query_head_groups_that_are_not_instantiated_fall_back():
    """Both kernels are instantiated for 1, 2, 4 or 8 query heads per KV head: a plan for 3 must say no (the engine then runs the
    per-kernel path) instead of saying yes and failing at the first launch."""
    ms = dataclasses.replace(PRESETS["tinyllama-1.1b"], heads=12, kv_heads=4, hidden=768, kv_hidden=256)
    for wb in (0, 1):
        ok, *_, why = plan(ms, BF16, BF16, word_based=wb)
        assert not ok and "head shape" in why


def test_word_based_kernel_plans():
    """The barrier-free kernel: on one GPU it takes the headline shape too (opt-in SLLM_ENGINE_MEGA_LL), with bf16 or int8 group-64
    tiles; its grid shrinks to the smallest phase's tile rows when a shard has fewer of them than SMs."""
    ms = PRESETS["llama2-7b"]
    ok, grid, nbytes, nsplit, why = plan(ms, BF16, BF16, word_based=1)
    assert ok and grid == B200_SMS and nbytes + 1024 <= B200_SMEM, why
    ok, grid, nbytes, nsplit, why = plan(ms, INT8, BF16, word_based=1)
    assert ok and grid == B200_SMS, why
    tiny = PRESETS["tiny_gqa"]                       # 128 x 128 wo: 64 two-row units in tiles of four rows = 32 tile rows
    ok, grid, *_ = plan(tiny, F32, F32, word_based=1)
    assert ok and 1 <= grid < B200_SMS


@pytest.mark.parametrize("wd", [F32, BF16, INT8])
@pytest.mark.parametrize("rows,cols,kind", [
    (12288, 4096, 0), (4096, 4096, 1), (22016, 4096, 2), (4096, 11008, 3), (32000, 4096, 4),     # llama2-7b
    (6144, 4096, 0), (28672, 4096, 2), (4096, 14336, 3), (128256, 4096, 4),                     # llama3-8b
    (2560, 2048, 0), (2048, 5632, 3), (864, 288, 0), (288, 768, 3), (300, 96, 4), (301, 512, 1), (1376, 512, 3)])
def test_tile_geometry_invariants(rows, cols, kind, wd):
    if cols % E[wd] or (wd == INT8 and cols % 64):
        pytest.skip("row length not a whole number of 16-byte chunks / int8 groups")
    nchunks = cols // E[wd]
    if nchunks * 16 > 32 * 1024:                                  # rows longer than 32 KB: 16 slices of more than 2 KB, no tile fits a slot
        with pytest.raises(_lib.SllmError) as ei:
            geometry(rows, cols, kind, wd)
        assert ei.value.code == _lib.ENOTSUP
        return
    ks, sc, r, ntr, tile_bytes, matrix_bytes = geometry(rows, cols, kind, wd)
    assert ks in (1, 2, 4, 8, 16) and r in (2, 4)
    assert ks * sc >= nchunks and (ks == 1 or (ks // 2) * sc < nchunks + (ks // 2) * 4)   # slices cover the row; not more slices than needed
    if ks < 16:
        assert sc * 16 <= 1024 + (48 if wd == INT8 else 0)      # a slice of a row is at most 1 KB (int8: rounded up to whole groups)
    srow = tile_bytes // r - sc * 16                            # bytes of group scales per row inside a tile
    assert (srow == 0) == (wd != INT8)
    if wd == INT8:
        assert sc % 4 == 0 and srow % 16 == 0 and srow >= sc      # one fp32 scale per 4 chunks (64 weights)
    assert tile_bytes <= SLOT_BYTES and tile_bytes % 16 == 0      # one tile = one bulk copy into one ring slot
    phys = rows if kind == 2 else (rows + 1) // 2 * 2             # [up; gate] pairs (u, I + u): rows is even by construction
    assert ntr == -(-phys // r)
    assert matrix_bytes == ntr * ks * tile_bytes
    assert matrix_bytes >= rows * cols * (4 if wd == F32 else 2 if wd == BF16 else 1)   # padding only ever adds
    assert WARPS % ks == 0                                        # 16 / KS row groups stream concurrently


def test_tile_geometry_of_the_headline_matrices():
    """llama2-7b bf16: rows of 4096 columns are cut into 8 slices of 64 chunks (1 KB), four rows per 4 KB tile; the down
    projection's rows of 11008 columns into 16 slices of 86 chunks, two rows per tile."""
    assert geometry(12288, 4096, 0, BF16)[:3] == (8, 64, 4)
    ks, sc, r, ntr, tile_bytes, _ = geometry(4096, 11008, 3, BF16)
    assert (ks, sc, r) == (16, 86, 2) and tile_bytes == 2 * 86 * 16 and ntr == 2048


def test_tile_geometry_rejects_bad_arguments():
    lib = _lib.load()
    assert lib.sllm_mega_tile_geometry(0, 4096, 0, BF16, None, None, None, None, None, None) != 0
    assert lib.sllm_mega_tile_geometry(16, 4100, 0, BF16, None, None, None, None, None, None) != 0     # not whole 16-byte chunks
    assert lib.sllm_mega_tile_geometry(16, 4096, 7, BF16, None, None, None, None, None, None) != 0
