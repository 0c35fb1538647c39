"""ctypes loader for lib/libsllm_b200.so (the C ABI declared in include/sllm_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises. The library is built in-tree by
``__graft_entry__.build()`` / ``make -C simplellminference_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLLM_LIB") or os.path.join(HERE, "lib", "libsllm_b200.so")


class SllmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsllm_b200 error {code}: {msg}")
        self.code = code


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("vocab", "head_dim", "hidden", "kv_hidden", "inter", "max_len", "layers", "heads", "kv_heads")] + \
               [("eps", C.c_float), ("theta", C.c_float)]


class EngineConfig(C.Structure):
    _fields_ = [("shape", Shape), ("w_dtype", C.c_int32), ("kv_dtype", C.c_int32), ("group", C.c_int32),
                ("tp_rank", C.c_int32), ("tp_size", C.c_int32), ("flags", C.c_int32)]


ENGINE_UNFUSED, ENGINE_NO_GRAPH, ENGINE_PDL, ENGINE_P2P_ALLREDUCE, ENGINE_MEGAKERNEL, ENGINE_MEGA_LL = 1, 2, 4, 8, 16, 32
ENGINE_MEGA_FUSE_DOWN = 64
ENGINE_MEGA_V2 = 128
EINVAL, ENOTSUP, ENOMEM, ESTATE, ECOMM = -1, -2, -3, -4, -5   # SLLM_E* of include/sllm_b200.h

_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64
_F = C.c_float

# name -> (restype, argtypes); every symbol include/sllm_b200.h declares
SIGNATURES = {
    "sllm_last_error": (C.c_char_p, []),
    "sllm_abi_version": (C.c_int, []),
    "sllm_tune": (C.c_int, [_I, _I]),
    "sllm_device_info": (C.c_int, [C.POINTER(_I), C.POINTER(_I), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "sllm_add_f32": (C.c_int, [_P, _P, _P, _I, _P]),
    "sllm_embedding": (C.c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _P]),
    "sllm_rmsnorm_f32": (C.c_int, [_P, _P, _P, _I, _F, _P]),
    "sllm_gemv": (C.c_int, [_P, _P, _I, _P, _I, _P, _I, _I, _F, _P]),
    "sllm_rope_tables": (C.c_int, [_I, _I, _F, _P, _P, _P]),
    "sllm_rope_f32": (C.c_int, [_P, _P, _P, _I, _P, _P, _I, _I, _I, _P]),
    "sllm_mha_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "sllm_mha_decode": (C.c_int, [_P, _P, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "sllm_swiglu_f32": (C.c_int, [_P, _P, _P, _I, _P]),
    "sllm_argmax_f32": (C.c_int, [_P, _I, _P, _P]),
    "sllm_store_kv_row": (C.c_int, [_P, _P, _I, _I, _P]),
    "sllm_sample_f32": (C.c_int, [_P, _I, _F, _I, _F, C.c_uint64, C.c_uint64, _P, _P]),
    "sllm_synth_fill": (C.c_int, [C.POINTER(Shape), C.c_uint64, _I, _L, _L, _L, _L, _L, _P, _I, _P, _I, _P]),
    "sllm_convert_weights": (C.c_int, [_P, _P, _I, _P, _I, _L, _L, _P]),
    "sllm_engine_create": (C.c_int, [C.POINTER(EngineConfig), _P, C.POINTER(_P)]),
    "sllm_engine_destroy": (None, [_P]),
    "sllm_engine_load_synthetic": (C.c_int, [_P, C.c_uint64]),
    "sllm_engine_load_blob_f32": (C.c_int, [_P, _P, _L]),
    "sllm_comm_unique_id": (C.c_int, [_P]),
    "sllm_engine_init_comm": (C.c_int, [_P, _P]),
    "sllm_engine_p2p_export": (C.c_int, [_P, _P]),
    "sllm_engine_p2p_import": (C.c_int, [_P, _P]),
    "sllm_engine_prefill_p2p_export": (C.c_int, [_P, _P]),
    "sllm_engine_prefill_p2p_import": (C.c_int, [_P, _P]),
    "sllm_engine_forward": (C.c_int, [_P, _I, _I, _P, _P]),
    "sllm_engine_greedy": (C.c_int, [_P, _P, _I, _I, _P]),
    "sllm_engine_set_state": (C.c_int, [_P, _I, _I]),
    "sllm_engine_enqueue_steps": (C.c_int, [_P, _I]),
    "sllm_engine_read_tokens": (C.c_int, [_P, _P, _I]),
    "sllm_engine_prefill": (C.c_int, [_P, _P, _I, _I]),
    "sllm_engine_prefill_supported": (C.c_int, [_P]),
    "sllm_prefill_gemm_bf16": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "sllm_prefill_gemm_plan": (C.c_int, [_I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "sllm_mega_plan": (C.c_int, [C.POINTER(Shape), _I, _I, _I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_L), C.POINTER(_I)]),
    "sllm_mega_repack_down_t": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    "sllm_mega_tile_geometry": (C.c_int, [_I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_L)]),
    "sllm_prefill_attention": (C.c_int, [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "sllm_kvpages_create": (_P, [_I, _I, _I, _I]),
    "sllm_kvpages_destroy": (None, [_P]),
    "sllm_kvpages_reserve": (_I, [_P, _I, _I]),
    "sllm_kvpages_release": (C.c_int, [_P, _I]),
    "sllm_kvpages_free_count": (_I, [_P]),
    "sllm_kvpages_held": (_I, [_P, _I]),
    "sllm_kvpages_table": (C.POINTER(_I), [_P]),
    "sllm_batch_create": (C.c_int, [_P, _I, _I, _I, _I, C.POINTER(_P)]),
    "sllm_batch_destroy": (None, [_P]),
    "sllm_batch_add": (C.c_int, [_P, _P, _I, C.POINTER(_I)]),
    "sllm_batch_remove": (C.c_int, [_P, _I]),
    "sllm_batch_set_sampling": (C.c_int, [_P, _I, _F, _I, _F, C.c_uint64]),
    "sllm_batch_step": (C.c_int, [_P, _I]),
    "sllm_batch_set_tensor_cores": (C.c_int, [_P, _I]),
    "sllm_batch_read": (C.c_int, [_P, _I, _P, _I, C.POINTER(_I)]),
    "sllm_batch_logits": (C.c_int, [_P, _I, _P]),
    "sllm_batch_buffer": (C.c_int, [_P, _I, C.POINTER(_P), C.POINTER(_L), C.POINTER(_I)]),
    "sllm_batch_free_pages": (_I, [_P]),
    "sllm_batch_position": (_I, [_P, _I]),
    "sllm_batch_step_bytes": (_L, [_P]),
    "sllm_batch_total_launches": (_L, [_P]),
    "sllm_batch_arena_bytes": (_L, [C.POINTER(Shape), _I, _I, _I, _I]),
    "sllm_engine_buffer": (C.c_int, [_P, _I, C.POINTER(_P), C.POINTER(_L), C.POINTER(_I)]),
    "sllm_engine_step_bytes": (_L, [_P, _I]),
    "sllm_engine_enqueue_kernel": (C.c_int, [_P, _I, _I]),
    "sllm_engine_kernel_bytes": (_L, [_P, _I, _I]),
    "sllm_engine_kv_layout": (_I, [_P]),
    "sllm_engine_mode": (C.c_char_p, [_P]),
    "sllm_engine_calibrate": (C.c_int, [_P, _I]),
    "sllm_engine_calibration": (C.c_float, [_P, _I]),
    "sllm_engine_step_launches": (_I, [_P]),
    "sllm_engine_total_launches": (_L, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library (once). Raises FileNotFoundError if it was not built — no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). simplellminference_b200 has no CPU or PyTorch fallback.")
        # libsllm_b200.so needs libnccl.so.2. torch bundles a newer NCCL than the system one under the same
        # soname; whichever is loaded first wins for the whole process, so let torch (which needs the newer
        # one) load its copy first. A pure C/C++ host simply links the system NCCL.
        import torch  # noqa: F401
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise SllmError(rc, load().sllm_last_error().decode(errors="replace"))
