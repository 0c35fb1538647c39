"""Decode engine: Python face of sllm_engine_* (include/sllm_b200.h), the replacement for
model::LlamaModel::{init, forward, predict} (reference source/model/model.cpp:22-187)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .config import ModelShape, F32, BF16, INT8

_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16, INT8: torch.int8}
# ModelBufferType numbering of the reference (include/model/model.h:14-34)
BUF = dict(key_cache=2, value_cache=3, emb_output=4, rms_output=5, query=6, mha_output=8, att_output=9, ffn_input=10,
           up_output=11, gate_output=12, swi_output=14, ffn_output=15, model_pred=16, sin_cache=17, cos_cache=18)


class Engine:
    def __init__(self, shape: ModelShape, w_dtype: int = BF16, kv_dtype: int = BF16, group: int = 64, tp_rank: int = 0,
                 tp_size: int = 1, fused: bool = True, graph: bool = True, pdl: bool = False, mega: bool = False, mega_ll: bool = False, p2p_allreduce: bool = False,
                 stream: torch.cuda.Stream | None = None, mega_fuse_down: bool = False, mega_v2: bool = False):
        self.lib = _lib.load()
        self.shape, self.w_dtype, self.kv_dtype, self.group = shape, w_dtype, kv_dtype, group
        self.tp_rank, self.tp_size = tp_rank, tp_size
        flags = (0 if fused else _lib.ENGINE_UNFUSED) | (0 if graph else _lib.ENGINE_NO_GRAPH) | \
                (_lib.ENGINE_PDL if pdl else 0) | (_lib.ENGINE_MEGAKERNEL if mega else 0) | (_lib.ENGINE_MEGA_LL if mega_ll else 0) | (_lib.ENGINE_P2P_ALLREDUCE if p2p_allreduce else 0) | \
                (_lib.ENGINE_MEGA_FUSE_DOWN if mega_fuse_down else 0) | (_lib.ENGINE_MEGA_V2 if mega_v2 else 0)   # see SLLM_ENGINE_MEGA_* in include/sllm_b200.h
        cfg = _lib.EngineConfig(_lib.Shape(shape.vocab, shape.head_dim, shape.hidden, shape.kv_hidden, shape.inter, shape.max_len,
                                           shape.layers, shape.heads, shape.kv_heads, shape.eps, shape.theta),
                                w_dtype, kv_dtype, group, tp_rank, tp_size, flags)
        # no stream given -> the engine owns one (the legacy default stream cannot be graph-captured); every
        # host-visible call below ends with a synchronise of that stream, so torch reads that follow are safe
        self.stream = stream
        h = C.c_void_p()
        _lib.check(self.lib.sllm_engine_create(C.byref(cfg), stream.cuda_stream if stream is not None else None, C.byref(h)))
        self.h = h
        self._pin_tokens = None

    # -- weights --
    def load_synthetic(self, seed: int = 1234) -> "Engine":
        _lib.check(self.lib.sllm_engine_load_synthetic(self.h, seed))
        return self

    def load_blob(self, blob: np.ndarray) -> "Engine":
        """blob: host fp32 array in the reference's tensor order (model.cpp:340-468)."""
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        _lib.check(self.lib.sllm_engine_load_blob_f32(self.h, blob.ctypes.data, blob.size))
        return self

    def calibrate(self, rounds: int = 2) -> "Engine":
        """Size every CTA's share of the persistent kernel's phases by its measured HBM streaming rate (sllm_engine_calibrate).
        Call after loading weights and before use (it runs a few decode steps from position 0 and resets the step state)."""
        _lib.check(self.lib.sllm_engine_calibrate(self.h, rounds))
        return self

    def calibration(self) -> np.ndarray:
        """Relative time per byte of every CTA as last measured (1.0 = mean); empty if never calibrated."""
        out = []
        while True:
            v = float(self.lib.sllm_engine_calibration(self.h, len(out)))
            if v == 0.0:
                break
            out.append(v)
        return np.asarray(out, dtype=np.float32)

    # -- tensor parallel bootstrap (torch.distributed is only the messenger of the 128-byte NCCL id) --
    def init_comm(self, dist=None) -> "Engine":
        import torch.distributed as td
        dist = dist or td
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.tp_rank == 0:
            buf = (C.c_uint8 * 128)()
            _lib.check(self.lib.sllm_comm_unique_id(buf))
            ident = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            ident = ident.cuda()
        dist.broadcast(ident, src=0)
        raw = bytes(ident.cpu().numpy().tobytes())
        _lib.check(self.lib.sllm_engine_init_comm(self.h, raw))
        return self

    def init_p2p(self, dist=None) -> "Engine":
        """Peer-memory all-reduce (engine created with p2p_allreduce=True): exchange the 64-byte CUDA IPC handles of the
        ranks' receive areas; torch.distributed is only the messenger."""
        import torch.distributed as td
        dist = dist or td
        buf = (C.c_uint8 * 64)()
        _lib.check(self.lib.sllm_engine_p2p_export(self.h, buf))
        mine = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        allh = [torch.zeros_like(mine) for _ in range(self.tp_size)]
        dist.all_gather(allh, mine)
        raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        dist.barrier()
        _lib.check(self.lib.sllm_engine_p2p_import(self.h, raw))
        dist.barrier()
        # the prefill exchange blocks (csrc/prefill_tp.cu), when the engine has one: batched prefill then needs no communicator at all
        rc = self.lib.sllm_engine_prefill_p2p_export(self.h, buf)
        if rc == 0:
            mine = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
            if dist.get_backend() == "nccl":
                mine = mine.cuda()
            allh = [torch.zeros_like(mine) for _ in range(self.tp_size)]
            dist.all_gather(allh, mine)
            raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
            dist.barrier()
            _lib.check(self.lib.sllm_engine_prefill_p2p_import(self.h, raw))
            dist.barrier()
        self.prefill_p2p = rc == 0
        return self

    # -- LlamaModel::forward: one token at one position; returns (logits np.float32[V_local], next_token) --
    def forward(self, token: int, pos: int, want_logits: bool = True):
        n_loc = self.shape.vocab // self.tp_size
        logits = np.empty(n_loc, np.float32) if want_logits else None
        nxt = C.c_int32(-1)
        _lib.check(self.lib.sllm_engine_forward(self.h, token, pos, logits.ctypes.data if want_logits else None, C.byref(nxt)))
        return logits, nxt.value

    # -- greedy loop of LlamaModel::predict: returns the n_total-1 tokens that follow prompt[0] --
    def greedy(self, prompt, n_total: int) -> np.ndarray:
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_total - 1, np.int32)
        _lib.check(self.lib.sllm_engine_greedy(self.h, prompt.ctypes.data, prompt.size, n_total, out.ctypes.data))
        return out

    # -- device-resident stepping (benchmarks) --
    def set_state(self, token: int, pos: int) -> None:
        _lib.check(self.lib.sllm_engine_set_state(self.h, token, pos))

    def enqueue_steps(self, n: int) -> None:
        _lib.check(self.lib.sllm_engine_enqueue_steps(self.h, n))

    def read_tokens(self, n: int) -> np.ndarray:
        out = np.empty(n, np.int32)
        _lib.check(self.lib.sllm_engine_read_tokens(self.h, out.ctypes.data, n))
        return out

    def prefill(self, prompt, start_pos: int = 0) -> None:
        """Batched prompt processing on the tensor cores (sllm_engine_prefill): KV cache filled for the whole prompt, the
        last prompt token's logits in model_pred, its arg-max as the current token. Asynchronous on the engine's stream."""
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        _lib.check(self.lib.sllm_engine_prefill(self.h, prompt.ctypes.data, prompt.size, start_pos))

    @property
    def prefill_supported(self) -> bool:
        return bool(self.lib.sllm_engine_prefill_supported(self.h))

    # -- introspection --
    def buffer(self, name_or_id) -> torch.Tensor:
        """Device view (no copy) of a named buffer (reference ModelBufferType names) as a flat torch tensor."""
        bid = BUF[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        ptr, n, dt = C.c_void_p(), C.c_int64(), C.c_int32()
        _lib.check(self.lib.sllm_engine_buffer(self.h, bid, C.byref(ptr), C.byref(n), C.byref(dt)))
        tdt = _TORCH_DT[dt.value]
        # wrap device memory without copying: go through __cuda_array_interface__
        t = torch.as_tensor(_CudaView(ptr.value, n.value, tdt), device="cuda")
        return t.view(torch.bfloat16) if tdt == torch.bfloat16 else t

    KERNELS = dict(embed=0, qkv=1, mha=2, wo=3, gate_up=4, down=5, cls=6)

    def enqueue_kernel(self, kind: str, layer: int) -> None:
        _lib.check(self.lib.sllm_engine_enqueue_kernel(self.h, self.KERNELS[kind], layer))

    def kernel_bytes(self, kind: str, pos: int) -> int:
        return int(self.lib.sllm_engine_kernel_bytes(self.h, self.KERNELS[kind], pos))

    def kv_row(self, which: str, layer: int, pos: int) -> torch.Tensor:
        """Row (layer, pos) of the key or value cache in the REFERENCE's order [kv_heads*head_dim], whatever the
        engine's internal cache layout is (sllm_engine_kv_layout)."""
        s = self.shape
        kvh, hd = s.kv_heads // self.tp_size, s.head_dim
        buf = self.buffer("key_cache" if which == "k" else "value_cache")
        if self.lib.sllm_engine_kv_layout(self.h) == 1:
            return buf.view(s.layers, kvh, s.max_len, hd)[layer, :, pos, :].reshape(-1)
        return buf.view(s.layers, s.max_len, kvh * hd)[layer, pos]

    def step_bytes(self, pos: int) -> int:
        return int(self.lib.sllm_engine_step_bytes(self.h, pos))

    @property
    def mode(self) -> str:
        return self.lib.sllm_engine_mode(self.h).decode()

    @property
    def step_launches(self) -> int:
        return int(self.lib.sllm_engine_step_launches(self.h))

    @property
    def total_launches(self) -> int:
        return int(self.lib.sllm_engine_total_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.sllm_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaView:
    """Minimal __cuda_array_interface__ carrier so torch can view engine-owned device memory."""

    _TYPESTR = {torch.float32: "<f4", torch.bfloat16: "<i2", torch.int8: "|i1", torch.int32: "<i4"}

    def __init__(self, ptr: int, n: int, dtype):
        self._dtype = dtype
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": self._TYPESTR[dtype], "data": (ptr, False), "version": 3}
