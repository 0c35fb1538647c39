"""Batched multi-sequence decode over a paged KV cache: Python face of sllm_batch_* / sllm_kvpages_*
(include/sllm_b200.h). Additive to the reference, which decodes one sequence at a time (include/model/model.h:15-18,
source/model/model.cpp:148-185); every sequence of a batch follows the semantics of ``Engine.greedy`` — token for token with the default
fp32-activation step; within the bf16-operand tolerance of the prefill with the opt-in tensor-core step (``set_tensor_cores``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .config import BF16
from .engine import Engine


class KvPages:
    """Host bookkeeping of the paged cache (free stack + per-sequence page lists); no device involved."""

    def __init__(self, n_pages: int, page_len: int, max_seqs: int, max_pages_per_seq: int):
        self.lib = _lib.load()
        self.h = self.lib.sllm_kvpages_create(n_pages, page_len, max_seqs, max_pages_per_seq)
        if not self.h:
            raise _lib.SllmError(_lib.EINVAL, self.lib.sllm_last_error().decode(errors="replace"))
        self.max_seqs, self.max_pages = max_seqs, max_pages_per_seq

    def reserve(self, seq: int, n_positions: int) -> int:
        """Cover positions [0, n_positions) of ``seq``; returns the number of pages newly taken. All or nothing."""
        rc = self.lib.sllm_kvpages_reserve(self.h, seq, n_positions)
        if rc < 0:
            _lib.check(rc)
        return rc

    def release(self, seq: int) -> None:
        _lib.check(self.lib.sllm_kvpages_release(self.h, seq))

    @property
    def free(self) -> int:
        return int(self.lib.sllm_kvpages_free_count(self.h))

    def held(self, seq: int) -> int:
        return int(self.lib.sllm_kvpages_held(self.h, seq))

    def table(self) -> np.ndarray:
        ptr = self.lib.sllm_kvpages_table(self.h)
        return np.ctypeslib.as_array(ptr, shape=(self.max_seqs, self.max_pages)).copy()

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.sllm_kvpages_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def arena_bytes(shape, max_seqs: int, page_len: int, n_pages: int, kv_dtype: int = BF16) -> int:
    """Device bytes a ``BatchDecoder`` of these parameters takes on an engine of ``shape`` (host arithmetic, no GPU)."""
    lib = _lib.load()
    cs = _lib.Shape(shape.vocab, shape.head_dim, shape.hidden, shape.kv_hidden, shape.inter, shape.max_len, shape.layers, shape.heads,
                    shape.kv_heads, shape.eps, shape.theta)
    n = int(lib.sllm_batch_arena_bytes(C.byref(cs), max_seqs, page_len, n_pages, kv_dtype))
    if n < 0:
        raise _lib.SllmError(_lib.EINVAL, lib.sllm_last_error().decode(errors="replace"))
    return n


def pages_that_fit(shape, max_seqs: int, page_len: int, hbm_bytes: int, kv_dtype: int = BF16) -> int:
    """Largest page pool (``n_pages``) whose decoder fits ``hbm_bytes`` of device memory, capped at what ``max_seqs`` sequences of
    ``shape.max_len`` positions can ever hold; 0 if not even one page per slot fits. The pools grow linearly in ``n_pages``."""
    cap = max_seqs * -(-shape.max_len // page_len)
    fixed = arena_bytes(shape, max_seqs, page_len, 1, kv_dtype)
    per_page = 2 * shape.layers * shape.kv_heads * page_len * shape.head_dim * (4 if kv_dtype != BF16 else 2)
    n = min(cap, int(max(0, hbm_bytes - fixed) // per_page) + 1)
    while n > 0 and arena_bytes(shape, max_seqs, page_len, n, kv_dtype) > hbm_bytes:   # the 1 MiB / 256-byte round-ups
        n -= 1
    return n if n >= max_seqs else 0


class BatchDecoder:
    """Up to ``max_seqs`` sequences stepping together over the weights of ``engine`` (one GPU, not a megakernel engine)."""

    def __init__(self, engine: Engine, max_seqs: int, page_len: int = 64, n_pages: int | None = None, kv_dtype: int = BF16,
                 tensor_cores: bool = False):
        self.lib = _lib.load()
        self.engine = engine   # keeps the weights alive
        if n_pages is None:    # enough for every slot to reach the engine's max_len ...
            n_pages = max_seqs * ((engine.shape.max_len + page_len - 1) // page_len)
            if arena_bytes(engine.shape, max_seqs, page_len, n_pages, kv_dtype) > (8 << 30):
                # ... unless that is a large pool (64 Llama-2-7B sequences of 4096 positions = 128 GiB): then as many pages as 90 % of the
                # free HBM carries (sllm_device_info), never fewer than one per slot
                free = C.c_size_t(0)
                if self.lib.sllm_device_info(None, None, None, C.byref(free)) == 0 and free.value:
                    fit = pages_that_fit(engine.shape, max_seqs, page_len, int(free.value * 0.9), kv_dtype)
                    if fit:
                        n_pages = min(n_pages, fit)
        h = C.c_void_p()
        _lib.check(self.lib.sllm_batch_create(engine.h, max_seqs, page_len, n_pages, kv_dtype, C.byref(h)))
        self.h = h
        self.max_seqs, self.page_len, self.n_pages, self.kv_dtype = max_seqs, page_len, n_pages, kv_dtype
        self.max_len = engine.shape.max_len   # positions one sequence can reach (sllm_batch_step refuses to go past it)
        if tensor_cores:
            self.set_tensor_cores(True)

    def add(self, prompt) -> int:
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        slot = C.c_int32(-1)
        _lib.check(self.lib.sllm_batch_add(self.h, prompt.ctypes.data, prompt.size, C.byref(slot)))
        return slot.value

    def remove(self, slot: int) -> None:
        _lib.check(self.lib.sllm_batch_remove(self.h, slot))

    def set_sampling(self, slot: int, temperature: float, top_k: int = 0, top_p: float = 0.0, seed: int = 0) -> None:
        """Draw this sequence's tokens (kernels.sample semantics, keyed by (seed, position)) instead of arg-max; temperature <= 0 = arg-max."""
        _lib.check(self.lib.sllm_batch_set_sampling(self.h, slot, float(temperature), int(top_k), float(top_p), int(seed)))

    def set_tensor_cores(self, on: bool = True) -> None:
        """Opt-in: the projections of every step as tcgen05 GEMMs over the live sequences' rows (weights read once per step however many
        sequences are live; bf16 GEMM operands, so results match the reference within the prefill's tolerance, not bit for bit)."""
        _lib.check(self.lib.sllm_batch_set_tensor_cores(self.h, int(bool(on))))

    def step(self, n_steps: int = 1) -> None:
        _lib.check(self.lib.sllm_batch_step(self.h, n_steps))

    def tokens(self, slot: int) -> np.ndarray:
        """The tokens that followed positions 0.. of the slot's sequence (what ``Engine.greedy`` returns)."""
        n = self.position(slot)
        out = np.empty(max(n, 1), np.int32)
        got = C.c_int32(0)
        _lib.check(self.lib.sllm_batch_read(self.h, slot, out.ctypes.data, max(n, 0), C.byref(got)))
        return out[:got.value]

    def logits(self, slot: int) -> np.ndarray:
        out = np.empty(self.engine.shape.vocab, np.float32)
        _lib.check(self.lib.sllm_batch_logits(self.h, slot, out.ctypes.data))
        return out

    def buffer(self, name_or_id):
        """Device view (no copy) of a per-slot buffer as a torch tensor [max_seqs, width] (reference ModelBufferType names);
        the page pools ("key_cache" / "value_cache") come back as [pages, layers, kv_heads, page_len, head_dim]."""
        import torch
        from .engine import BUF, _CudaView, _TORCH_DT
        bid = BUF[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        ptr, n, dt = C.c_void_p(), C.c_int64(), C.c_int32()
        _lib.check(self.lib.sllm_batch_buffer(self.h, bid, C.byref(ptr), C.byref(n), C.byref(dt)))
        tdt = _TORCH_DT[dt.value]
        t = torch.as_tensor(_CudaView(ptr.value, n.value, tdt), device="cuda")
        t = t.view(torch.bfloat16) if tdt == torch.bfloat16 else t
        if bid in (2, 3):
            s = self.engine.shape
            return t.view(self.n_pages, s.layers, s.kv_heads, self.page_len, s.head_dim)
        return t.view(self.max_seqs, -1)

    def position(self, slot: int) -> int:
        return int(self.lib.sllm_batch_position(self.h, slot))

    @property
    def free_pages(self) -> int:
        return int(self.lib.sllm_batch_free_pages(self.h))

    def step_bytes(self) -> int:
        return int(self.lib.sllm_batch_step_bytes(self.h))

    @property
    def total_launches(self) -> int:
        return int(self.lib.sllm_batch_total_launches(self.h))

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.sllm_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
