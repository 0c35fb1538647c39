"""simplellminference_b200 — B200-native (sm_100a) forward hot path behind the SimpleLLMInference op/kernel API.

Layers (bottom-up):
  csrc/        hand-written CUDA kernels + the extern-"C" boundary (include/sllm_b200.h) -> lib/libsllm_b200.so
  host/        C++ mirror of the reference's mem:: / op:: / kernel:: / model:: interfaces over that C ABI
  _lib.py      ctypes loader (fails loudly when the CUDA library is missing: there is no CPU fallback)
  kernels.py   one Python wrapper per op launcher (torch tensors are only the carriers of device pointers)
  engine.py    the decode engine (arena + KV cache + CUDA-graph decode step, tensor parallel)
  batch.py     batched multi-sequence decode over a paged KV cache (sllm_batch_* / sllm_kvpages_*)
  config.py    the model-shape presets named by BASELINE.json
"""
from .config import ModelShape, PRESETS, F32, BF16, INT8  # noqa: F401

__all__ = ["ModelShape", "PRESETS", "F32", "BF16", "INT8"]
