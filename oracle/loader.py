"""oracle/loader.py — TEST INFRASTRUCTURE ONLY.

ctypes front-end for the two CPU oracles:

* ``port``  — oracle/_build/liboracle_port.so, the plain-C restatement (oracle/llama_oracle.c). Always buildable.
* ``ref``   — oracle/_ref/libref_oracle.so, the UNMODIFIED reference CPU path compiled from /root/reference by
              oracle/Makefile (``ref_fast`` = same sources at -O3 -march=x86-64-v3, speed baseline only).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import
this module; the product package ``simplellminference_b200`` never does (tests/test_no_oracle_in_product.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "liboracle_port.so")
PORT_FMA_SO = os.path.join(HERE, "_build", "liboracle_port_fma.so")   # yardstick build (FMA contraction), see oracle/Makefile
PORT_FAST_SO = os.path.join(HERE, "_build", "liboracle_port_fast.so")  # yardstick build (-Ofast: re-associated sums)
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libref_oracle_fast.so")

F32, BF16, INT8 = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


class SynShape(C.Structure):
    """Mirror of ``syn_shape`` (oracle/synth_weights.h) == the reference's LlamaModelConfig fields."""

    _fields_ = [(n, C.c_int32) for n in
                ("vocab", "head_dim", "hidden", "kv_hidden", "inter", "max_len", "layers", "heads", "kv_heads")] + \
               [("eps", C.c_float), ("theta", C.c_float)]


@dataclass(frozen=True)
class Shape:
    vocab: int
    head_dim: int
    hidden: int
    kv_hidden: int
    inter: int
    max_len: int
    layers: int
    heads: int
    kv_heads: int
    eps: float = 1e-5
    theta: float = 10000.0

    def c(self) -> SynShape:
        return SynShape(self.vocab, self.head_dim, self.hidden, self.kv_hidden, self.inter, self.max_len,
                        self.layers, self.heads, self.kv_heads, self.eps, self.theta)

    def cfg9(self) -> np.ndarray:
        return np.array([self.vocab, self.head_dim, self.hidden, self.kv_hidden, self.inter, self.max_len,
                         self.layers, self.heads, self.kv_heads], dtype=np.int32)


def build(which: str = "port") -> None:
    """(Re)build an oracle library with oracle/Makefile. ``ref`` needs /root/reference."""
    subprocess.run(["make", "-s", "-C", HERE, which], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def cpu_supports_v3() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return all(f in flags for f in (" avx2", " fma", " bmi2"))


# ----------------------------------------------------------------------------------------------- port ------
class Port:
    """The plain-C restatement + the synthetic weight generator."""

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build("port")
        L = self.lib = C.CDLL(path)
        SP = C.POINTER(SynShape)
        L.syn_blob_floats.restype = C.c_int64
        L.syn_blob_floats.argtypes = [SP]
        L.syn_segment.argtypes = [SP, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.syn_value.restype = C.c_float
        L.syn_value.argtypes = [C.c_uint64, C.c_int, C.c_int64, C.c_float, C.c_float]
        L.syn_round_bf16.restype = C.c_float
        L.syn_round_bf16.argtypes = [C.c_float]
        L.syn_fill_segment.argtypes = [SP, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, _f32p, C.c_int]
        L.syn_fill_blob.argtypes = [SP, C.c_uint64, C.c_int, C.c_int, _f32p, C.c_int]
        L.syn_fill_segment_int8.argtypes = [SP, C.c_uint64, C.c_int, C.c_int, C.c_int64, C.c_int64, _i8p, _f32p]
        L.orc_embedding.argtypes = [C.c_int32, _f32p, _f32p, C.c_int32, C.c_int32]
        L.orc_rmsnorm.argtypes = [_f32p, _f32p, _f32p, C.c_int32, C.c_float]
        L.orc_matmul.argtypes = [_f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_float]
        L.orc_rope_cache.argtypes = [C.c_int32, C.c_int32, C.c_float, _f32p, _f32p]
        L.orc_rope.argtypes = [_f32p, _f32p, C.c_int32, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32]
        L.orc_softmax.argtypes = [_f32p, C.c_int32]
        L.orc_mha.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p] + [C.c_int32] * 6
        L.orc_add.argtypes = [_f32p, _f32p, _f32p, C.c_int32]
        L.orc_swiglu.argtypes = [_f32p, _f32p, _f32p, C.c_int32]
        L.orc_argmax.restype = C.c_int32
        L.orc_argmax.argtypes = [_f32p, C.c_int32]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [SP, _f32p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_kv_bf16.argtypes = [C.c_void_p, C.c_int]
        L.orc_forward.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_greedy.restype = C.c_int32
        L.orc_greedy.argtypes = [C.c_void_p, _i32p, C.c_int32, C.c_int32, _i32p, C.c_void_p]
        L.orc_read.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, _f32p]
        L.orc_write.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, _f32p]

    # -- synthetic weights --
    def blob_floats(self, shape: Shape) -> int:
        return int(self.lib.syn_blob_floats(C.byref(shape.c())))

    def segment(self, shape: Shape, t: int):
        off, cnt, row = C.c_int64(), C.c_int64(), C.c_int64()
        mean, c = C.c_float(), C.c_float()
        self.lib.syn_segment(C.byref(shape.c()), t, C.byref(off), C.byref(cnt), C.byref(row), C.byref(mean), C.byref(c))
        return off.value, cnt.value, row.value, mean.value, c.value

    def fill_blob(self, shape: Shape, seed: int = 1234, wdtype: int = F32, group: int = 64, threads: int = 0) -> np.ndarray:
        blob = np.empty(self.blob_floats(shape), dtype=np.float32)
        self.lib.syn_fill_blob(C.byref(shape.c()), seed, wdtype, group, blob, threads or (os.cpu_count() or 1))
        return blob

    def fill_segment(self, shape: Shape, t: int, first: int, count: int, seed: int = 1234, wdtype: int = F32,
                     group: int = 64) -> np.ndarray:
        out = np.empty(count, dtype=np.float32)
        self.lib.syn_fill_segment(C.byref(shape.c()), seed, t, wdtype, group, first, count, out, 1)
        return out

    def fill_segment_int8(self, shape: Shape, t: int, first: int, count: int, seed: int = 1234, group: int = 64):
        q = np.empty(count, dtype=np.int8)
        sc = np.empty(count // group, dtype=np.float32)
        self.lib.syn_fill_segment_int8(C.byref(shape.c()), seed, t, group, first, count, q, sc)
        return q, sc

    # -- ops (numpy in / numpy out) --
    def embedding(self, token, table):
        out = np.empty(table.shape[1], np.float32)
        self.lib.orc_embedding(token, table, out, table.shape[0], table.shape[1])
        return out

    def rmsnorm(self, x, w, eps):
        y = np.empty_like(x)
        self.lib.orc_rmsnorm(x, w, y, x.size, eps)
        return y

    def matmul(self, x, W, scale=1.0):
        y = np.empty(W.shape[0], np.float32)
        self.lib.orc_matmul(x, W, y, W.shape[0], W.shape[1], scale)
        return y

    def rope_cache(self, head_dim, max_len, theta):
        s = np.empty((max_len, head_dim // 2), np.float32)
        c = np.empty_like(s)
        self.lib.orc_rope_cache(head_dim, max_len, theta, s, c)
        return s, c

    def rope(self, q, k, pos, sin_c, cos_c, head_dim):
        q, k = q.copy(), k.copy()
        self.lib.orc_rope(q, k, pos, sin_c, cos_c, q.size, k.size, head_dim)
        return q, k

    def mha(self, q, kc, vc, layer, pos, head_dim, heads, kv_heads):
        L, S, kv = kc.shape
        score = np.zeros((max(heads, head_dim), S), np.float32)
        out = np.empty(heads * head_dim, np.float32)
        self.lib.orc_mha(q, score, kc, vc, out, layer, pos, S, head_dim, heads, kv_heads)
        return out

    def add(self, a, b):
        out = np.empty_like(a)
        self.lib.orc_add(a, b, out, a.size)
        return out

    def swiglu(self, up, gate):
        out = np.empty_like(up)
        self.lib.orc_swiglu(up, gate, out, up.size)
        return out

    def argmax(self, logits):
        return int(self.lib.orc_argmax(logits, logits.size))

    def model(self, shape: Shape, blob: np.ndarray, threads: int = 1, kv_bf16: bool = False) -> "PortModel":
        return PortModel(self, shape, blob, threads, kv_bf16)


class PortModel:
    def __init__(self, port: Port, shape: Shape, blob: np.ndarray, threads: int, kv_bf16: bool):
        self.lib, self.shape, self._blob = port.lib, shape, blob
        self.h = self.lib.orc_create(C.byref(shape.c()), blob)
        self.lib.orc_set_threads(self.h, threads)
        self.lib.orc_set_kv_bf16(self.h, int(kv_bf16))

    def forward(self, token: int, pos: int) -> np.ndarray:
        logits = np.empty(self.shape.vocab, np.float32)
        self.lib.orc_forward(self.h, token, pos, logits.ctypes.data)
        return logits

    def step(self, token: int, pos: int) -> None:
        self.lib.orc_forward(self.h, token, pos, None)

    def greedy(self, prompt, n_total: int):
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_total - 1, np.int32)
        logits = np.empty(self.shape.vocab, np.float32)
        n = self.lib.orc_greedy(self.h, prompt, prompt.size, n_total, out, logits.ctypes.data)
        return out[:n], logits

    def read(self, buffer_id: int, offset: int, n: int) -> np.ndarray:
        out = np.empty(n, np.float32)
        self.lib.orc_read(self.h, buffer_id, offset, n, out)
        return out

    def write(self, buffer_id: int, offset: int, values: np.ndarray) -> None:
        """Overwrite part of the key (2) / value (3) cache, layout [L][S][kv] — injects a synthetic history."""
        values = np.ascontiguousarray(values, dtype=np.float32)
        self.lib.orc_write(self.h, buffer_id, offset, values.size, values)

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    __del__ = close


# ------------------------------------------------------------------------------------------------ ref ------
class Ref:
    """The unmodified reference CPU path (oracle/_ref). Raises FileNotFoundError if it was never built."""

    def __init__(self, fast: bool = False):
        path = REF_FAST_SO if fast else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = C.CDLL(path)
        L.ref_build_flags.restype = C.c_char_p
        L.ref_model_create.restype = C.c_void_p
        L.ref_model_create.argtypes = [_i32p, C.c_float, C.c_float, _f32p]
        L.ref_model_destroy.argtypes = [C.c_void_p]
        L.ref_model_forward.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.ref_model_greedy.restype = C.c_int32
        L.ref_model_greedy.argtypes = [C.c_void_p, _i32p, C.c_int32, C.c_int32, _i32p, C.c_void_p]
        L.ref_model_read.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, _f32p]
        L.ref_op_add.argtypes = [_f32p, _f32p, _f32p, C.c_int32]
        L.ref_op_embedding.argtypes = [C.c_int32, _f32p, _f32p, C.c_int32, C.c_int32]
        L.ref_op_rmsnorm.argtypes = [_f32p, _f32p, _f32p, C.c_int32, C.c_float]
        L.ref_op_matmul.argtypes = [_f32p, _f32p, _f32p, C.c_int32, C.c_int32]
        L.ref_op_swiglu.argtypes = [_f32p, _f32p, _f32p, C.c_int32]
        L.ref_rope_cache.argtypes = [C.c_int32, C.c_int32, C.c_float, _f32p, _f32p]
        L.ref_op_rope.argtypes = [_f32p, _f32p, C.c_int32, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32]
        L.ref_op_mha.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p] + [C.c_int32] * 7
        L.ref_op_argmax.restype = C.c_int32
        L.ref_op_argmax.argtypes = [_f32p, C.c_int32]
        self.flags = L.ref_build_flags().decode()

    def embedding(self, token, table):
        out = np.empty(table.shape[1], np.float32)
        self.lib.ref_op_embedding(token, table, out, table.shape[0], table.shape[1])
        return out

    def rmsnorm(self, x, w, eps):
        y = np.empty_like(x)
        self.lib.ref_op_rmsnorm(x, w, y, x.size, eps)
        return y

    def matmul(self, x, W):
        y = np.empty(W.shape[0], np.float32)
        self.lib.ref_op_matmul(x, W, y, W.shape[0], W.shape[1])
        return y

    def rope_cache(self, head_dim, max_len, theta):
        s = np.empty((max_len, head_dim // 2), np.float32)
        c = np.empty_like(s)
        self.lib.ref_rope_cache(head_dim, max_len, theta, s, c)
        return s, c

    def rope(self, q, k, pos, sin_c, cos_c, head_dim):
        """k must have q.size elements (the reference rotates k over q's length, rope_kernel.cpp:27-38)."""
        q, k = q.copy(), k.copy()
        assert k.size >= q.size
        self.lib.ref_op_rope(q, k, pos, sin_c, cos_c, sin_c.shape[0], q.size, head_dim)
        return q, k

    def mha(self, q, kc, vc, layer, pos, head_dim, heads, kv_heads):
        L, S, kv = kc.shape
        score = np.zeros((max(heads, head_dim), S), np.float32)
        out = np.empty(heads * head_dim, np.float32)
        self.lib.ref_op_mha(q, score, kc, vc, out, layer, pos, L, S, head_dim, heads, kv_heads)
        return out

    def add(self, a, b):
        out = np.empty_like(a)
        self.lib.ref_op_add(a, b, out, a.size)
        return out

    def swiglu(self, up, gate):
        out = np.empty_like(up)
        self.lib.ref_op_swiglu(up, gate, out, up.size)
        return out

    def argmax(self, logits):
        return int(self.lib.ref_op_argmax(logits, logits.size))

    def model(self, shape: Shape, blob: np.ndarray) -> "RefModel":
        return RefModel(self, shape, blob)


class RefModel:
    def __init__(self, ref: Ref, shape: Shape, blob: np.ndarray):
        self.lib, self.shape, self._blob = ref.lib, shape, blob
        self._cwd_guard()
        self.h = self.lib.ref_model_create(shape.cfg9(), shape.eps, shape.theta, blob)

    @staticmethod
    def _cwd_guard():
        # LlamaModel::forward() opens "layer_outputs_cpu.txt" in the cwd on every call (model.cpp:42).
        pass

    def forward(self, token: int, pos: int) -> np.ndarray:
        logits = np.empty(self.shape.vocab, np.float32)
        self.lib.ref_model_forward(self.h, token, pos, logits.ctypes.data)
        return logits

    def step(self, token: int, pos: int) -> None:
        self.lib.ref_model_forward(self.h, token, pos, None)

    def greedy(self, prompt, n_total: int):
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_total - 1, np.int32)
        logits = np.empty(self.shape.vocab, np.float32)
        n = self.lib.ref_model_greedy(self.h, prompt, prompt.size, n_total, out, logits.ctypes.data)
        return out[:n], logits

    def read(self, buffer_id: int, offset: int, n: int) -> np.ndarray:
        out = np.empty(n, np.float32)
        self.lib.ref_model_read(self.h, buffer_id, offset, n, out)
        return out

    def close(self):
        if self.h:
            self.lib.ref_model_destroy(self.h)
            self.h = None

    __del__ = close
