"""One Python wrapper per op launcher of include/sllm_b200.h.

torch tensors are only the carriers of device pointers and of the current stream; every computation happens in
libsllm_b200.so. Arguments mirror the reference's kernel::*_cuda functions (include/kernel/cuda/*.cuh).
"""
from __future__ import annotations

import torch

from . import _lib
from .config import F32, BF16, INT8


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "expects a contiguous CUDA float32 tensor"
    return t


def add(a, b, out=None):
    """kernel::add_kernel_cuda(a, b, out, n)"""
    out = torch.empty_like(a) if out is None else out
    _lib.check(_lib.load().sllm_add_f32(_p(_f32(a)), _p(_f32(b)), _p(_f32(out)), a.numel(), _stream()))
    return out


def embedding(token, table, w_dtype=F32, scales=None, group=64, d=None, vocab=None):
    """kernel::emb_kernel_cuda(token, W, out, vocab, d); `token` is an int or a CUDA int32 tensor."""
    vocab = table.shape[0] if vocab is None else vocab
    d = table.shape[1] if d is None else d
    out = torch.empty(d, dtype=torch.float32, device=table.device)
    tok_dev, tok = (token.data_ptr(), 0) if isinstance(token, torch.Tensor) else (None, int(token))
    _lib.check(_lib.load().sllm_embedding(tok_dev, tok, _p(table), w_dtype, _p(scales), group, _p(out), vocab, d, _stream()))
    return out


def rmsnorm(x, w, eps, out=None):
    """kernel::rmsnorm_kernel_cuda(x, w, y, d, eps)"""
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().sllm_rmsnorm_f32(_p(_f32(x)), _p(_f32(w)), _p(out), x.numel(), float(eps), _stream()))
    return out


def matmul(x, W, rows, cols, w_dtype=F32, scales=None, group=64, scale=1.0, out=None):
    """kernel::matmul_kernel_cuda(x, W, y, dim0=rows, dim1=cols, scale)"""
    out = torch.empty(rows, dtype=torch.float32, device=x.device) if out is None else out
    _lib.check(_lib.load().sllm_gemv(_p(_f32(x)), _p(W), w_dtype, _p(scales), group, _p(out), rows, cols, float(scale), _stream()))
    return out


def rope_tables(head_dim, max_len, theta, device="cuda"):
    """kernel::rope_cache_cal_cuda(head_dim, max_len, sin, cos, theta)"""
    s = torch.empty(max_len, head_dim // 2, dtype=torch.float32, device=device)
    c = torch.empty_like(s)
    _lib.check(_lib.load().sllm_rope_tables(head_dim, max_len, float(theta), _p(s), _p(c), _stream()))
    return s, c


def rope(q, k, pos, sin_t, cos_t, head_dim):
    """kernel::rope_kernel_cuda(q, k, pos, sin, cos, ...): in place; `pos` int or CUDA int32 tensor."""
    pos_dev, p = (pos.data_ptr(), 0) if isinstance(pos, torch.Tensor) else (None, int(pos))
    _lib.check(_lib.load().sllm_rope_f32(_p(_f32(q)), _p(_f32(k)), pos_dev, p, _p(sin_t), _p(cos_t), q.numel(), k.numel(), head_dim, _stream()))
    return q, k


def mha_workspace(heads, head_dim, max_len, device="cuda"):
    n = _lib.load().sllm_mha_workspace_bytes(heads, head_dim, max_len)
    return torch.zeros(n, dtype=torch.uint8, device=device)


def mha(q, key_cache, value_cache, layer, pos, head_dim, heads, kv_heads, workspace=None, kv_dtype=None):
    """kernel::mha_kernel_cuda(q, score, Kc, Vc, out, layer, pos, S, hd, ...); caches [L][S][kv]."""
    L, S, kv = key_cache.shape
    if kv_dtype is None:
        kv_dtype = BF16 if key_cache.dtype == torch.bfloat16 else F32
    workspace = mha_workspace(heads, head_dim, S, q.device) if workspace is None else workspace
    out = torch.empty(heads * head_dim, dtype=torch.float32, device=q.device)
    pos_dev, p = (pos.data_ptr(), 0) if isinstance(pos, torch.Tensor) else (None, int(pos))
    _lib.check(_lib.load().sllm_mha_decode(_p(_f32(q)), _p(key_cache), _p(value_cache), kv_dtype, _p(out), _p(workspace), layer,
                                           pos_dev, p, S, head_dim, heads, kv_heads, _stream()))
    return out


def swiglu(up, gate, out=None):
    """kernel::swiglu_kernel_cuda(up, gate, out, n): sigmoid(gate) * up"""
    out = torch.empty_like(up) if out is None else out
    _lib.check(_lib.load().sllm_swiglu_f32(_p(_f32(up)), _p(_f32(gate)), _p(out), up.numel(), _stream()))
    return out


def argmax(logits):
    """op::argmaxLayer::forward on the device: CUDA int32 tensor holding the first-max index."""
    idx = torch.empty(1, dtype=torch.int32, device=logits.device)
    _lib.check(_lib.load().sllm_argmax_f32(_p(_f32(logits)), logits.numel(), _p(idx), _stream()))
    return idx


def sample(logits, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, seed: int = 0, step: int = 0):
    """One draw from softmax(logits / temperature) under top-k / top-p (sllm_sample_f32): CUDA int32 tensor with the index."""
    idx = torch.empty(1, dtype=torch.int32, device=logits.device)
    _lib.check(_lib.load().sllm_sample_f32(_p(_f32(logits)), logits.numel(), float(temperature), int(top_k), float(top_p), int(seed), int(step),
                                           _p(idx), _stream()))
    return idx


def convert_weights(w_f32, w_dtype, group=64):
    """fp32 [rows][cols] -> storage dtype (+ scales for int8)."""
    rows, cols = w_f32.shape
    tdt = {F32: torch.float32, BF16: torch.bfloat16, INT8: torch.int8}[w_dtype]
    dst = torch.empty(rows, cols, dtype=tdt, device=w_f32.device)
    scales = torch.empty(rows, cols // group, dtype=torch.float32, device=w_f32.device) if w_dtype == INT8 else None
    _lib.check(_lib.load().sllm_convert_weights(_p(_f32(w_f32)), _p(dst), w_dtype, _p(scales), group, rows, cols, _stream()))
    return dst, scales


def prefill_gemm(a_bf16: torch.Tensor, w_bf16: torch.Tensor, bn: int = 0) -> torch.Tensor:
    """C[T][N] fp32 = A[T][K] . W[N][K]^T on the tcgen05 tensor cores (bf16 operands, fp32 accumulate in TMEM)."""
    assert a_bf16.dtype == torch.bfloat16 and w_bf16.dtype == torch.bfloat16 and a_bf16.is_contiguous() and w_bf16.is_contiguous()
    T, K = a_bf16.shape
    N, K2 = w_bf16.shape
    assert K == K2
    out = torch.empty(T, N, dtype=torch.float32, device=a_bf16.device)
    _lib.check(_lib.load().sllm_prefill_gemm_bf16(_p(a_bf16), _p(w_bf16), _p(out), T, N, K, bn, _stream()))
    return out


def prefill_attention(q_bf16: torch.Tensor, key_cache: torch.Tensor, value_cache: torch.Tensor, pos0: int, heads: int, kv_heads: int) -> torch.Tensor:
    """Causal attention of T queries at positions pos0.. over a head-major cache [kv_heads][max_len][head_dim]."""
    T = q_bf16.shape[0]
    kvh, max_len, hd = key_cache.shape
    assert kvh == kv_heads and q_bf16.shape[1] == heads * hd and q_bf16.dtype == torch.bfloat16
    kvd = F32 if key_cache.dtype == torch.float32 else BF16
    out = torch.empty_like(q_bf16)
    _lib.check(_lib.load().sllm_prefill_attention(_p(q_bf16), _p(key_cache), _p(value_cache), kvd, _p(out), T, pos0, max_len, hd, heads, kv_heads, _stream()))
    return out
