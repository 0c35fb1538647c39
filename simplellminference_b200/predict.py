"""Text in, text out: the driver part of LlamaModel::predict (reference source/model/model.cpp:142-187) around the engine.

The reference tokenises with sentencepiece (op::SPELayer, include/op/encode.h:17-28, source/op/encode.cpp:5-27), feeds the prompt
token by token, then feeds back the arg-max and prints each piece — and never stops before max_length (no EOS handling).
Here:
  * `SPELayer` keeps the reference's three calls (encode / decode / GetVocabularySize) over the SAME third-party library
    (the `sentencepiece` Python package instead of the C++ one); nothing of the tokenizer is re-implemented;
  * `predict` feeds the prompt token by token through the decode step, exactly as the reference does (model.cpp:157-166: fp32
    activations, the parity path), and generates in chunks of device-resident steps (one host synchronisation per chunk), optionally
    STOPPING at an EOS id (additive: `eos_id=None` is the reference's behaviour) and streaming decoded text to a callback like the
    reference's std::cout. `batched_prefill=True` (opt-in) runs the prompt as ONE tensor-core pass instead (sllm_engine_prefill):
    ~200x faster on a 512-token prompt, but the operands of its GEMMs are rounded to bf16, so its KV rows and logits agree with the
    reference within the prefill tolerance (DESIGN.md "Tolerances"), not bit for bit — token identity with the reference is only
    contracted for the default.
Token ids, not text, are what the parity tests compare; with the default arguments this module adds no arithmetic to the path.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import numpy as np


class SPELayer:
    """op::SPELayer: a sentencepiece model behind encode / decode / GetVocabularySize (encode.h:17-28)."""

    def __init__(self, model_file: str):
        try:
            import sentencepiece as spm
        except ImportError as ex:   # the tokenizer stays third-party, exactly as in the reference
            raise RuntimeError("SPELayer needs the third-party `sentencepiece` package") from ex
        self.processor_ = spm.SentencePieceProcessor()
        if not self.processor_.Load(model_file):   # encode.cpp:7-10 throws std::runtime_error
            raise RuntimeError(f"sentencepiece could not load {model_file}")

    def encode(self, text: str) -> list:
        return list(self.processor_.EncodeAsIds(text))

    def decode(self, ids: Iterable[int]) -> str:
        # a model may have more vocabulary rows than the tokenizer has pieces (padding rows): such ids carry no text
        n = self.GetVocabularySize()
        return self.processor_.DecodeIds([int(i) for i in ids if 0 <= int(i) < n])

    def GetVocabularySize(self) -> int:   # noqa: N802  (the reference's spelling)
        return int(self.processor_.GetPieceSize())


def sample_ids(engine, prompt_ids, max_length: int, temperature: float = 1.0, top_k: int = 0, top_p: float = 0.0, seed: int = 0,
               eos_id: Optional[int] = None) -> np.ndarray:
    """Like predict_ids, but every generated token is DRAWN on the device (kernels.sample on the logits the engine leaves in
    model_pred) instead of arg-max. Host-driven: one forward(token, pos) + one sample + a 4-byte read per token. Single GPU (under
    tensor parallelism each rank only holds its slice of the logits). Reproducible: the draw at position p uses (seed, step = p)."""
    from . import kernels
    if engine.tp_size != 1:
        raise ValueError("sample_ids: sampling needs the full logits vector (tp_size == 1)")
    prompt_ids = np.ascontiguousarray(prompt_ids, dtype=np.int32)
    n = int(prompt_ids.size)
    if n < 1 or n > max_length or max_length >= engine.shape.max_len:
        raise ValueError("sample_ids: need 1 <= len(prompt) <= max_length < max_len")
    out = [int(t) for t in prompt_ids[1:]]
    for pos in range(n - 1):                                  # prompt: logits unused (model.cpp:159-165)
        engine.forward(int(prompt_ids[pos]), pos, want_logits=False)
    tok = int(prompt_ids[-1])
    logits = engine.buffer("model_pred")
    for pos in range(n - 1, max_length):
        engine.forward(tok, pos, want_logits=False)
        tok = int(kernels.sample(logits, temperature, top_k, top_p, seed=seed, step=pos).item())
        out.append(tok)
        if eos_id is not None and tok == eos_id:
            break
    return np.asarray(out, dtype=np.int32)


def predict_ids(engine, prompt_ids, max_length: int, eos_id: Optional[int] = None, chunk: int = 16, batched_prefill: bool = False,
                on_tokens: Optional[Callable[[np.ndarray], None]] = None) -> np.ndarray:
    """Greedy loop of LlamaModel::predict on token ids. Returns the tokens that follow prompt[0] (prompt echo, then generated
    tokens) — at most max_length of them, like the reference; fewer if eos_id is given and generated (the EOS is the last token).
    Without an EOS hit the engine ends at position len(result) and a caller can continue with engine.enqueue_steps; after an
    EOS hit it has run on to the end of that chunk (the surplus tokens are dropped here)."""
    prompt_ids = np.ascontiguousarray(prompt_ids, dtype=np.int32)
    n = int(prompt_ids.size)
    if n < 1:
        raise ValueError("predict: empty prompt")
    if max_length >= engine.shape.max_len:
        raise ValueError("predict: max_length must be below the configured context (KV cache size)")
    if n > max_length:
        raise ValueError("predict: the prompt is longer than max_length")
    out = []
    if batched_prefill and n > 1 and engine.prefill_supported:
        engine.prefill(prompt_ids)                       # KV cache for the prompt, first generated token, position n
        first = engine.read_tokens(n)                    # n-1 prompt tokens (echo) + the first generated one
    else:
        first = engine.greedy(prompt_ids, n + 1)         # token by token, model.cpp:157-166
    out.extend(int(t) for t in first)
    if on_tokens:
        on_tokens(first)
    done = eos_id is not None and int(first[-1]) == eos_id
    while not done and len(out) < max_length:
        k = min(chunk, max_length - len(out))
        engine.enqueue_steps(k)                          # device-resident: token and position never visit the host
        toks = engine.read_tokens(k)                     # one synchronisation per chunk
        if eos_id is not None and (toks == eos_id).any():
            toks = toks[: int(np.flatnonzero(toks == eos_id)[0]) + 1]
            done = True
        out.extend(int(t) for t in toks)
        if on_tokens:
            on_tokens(toks)
    return np.asarray(out, dtype=np.int32)


def predict(engine, tokenizer: SPELayer, prompt: str, max_length: int, eos_id: Optional[int] = None, stream: Optional[Callable[[str], None]] = None,
            **kw) -> str:
    """LlamaModel::predict(prompt, max_length): returns the decoded text of prompt + continuation; `stream` receives each decoded
    chunk as it is produced (the reference prints piece by piece, model.cpp:155-182)."""
    if tokenizer.GetVocabularySize() > engine.shape.vocab:
        raise ValueError(f"tokenizer has {tokenizer.GetVocabularySize()} pieces, the model {engine.shape.vocab} rows")
    ids = tokenizer.encode(prompt)
    if not ids:
        raise ValueError("predict: the prompt encodes to no tokens")
    pieces = [ids[0]]
    if stream:
        stream(tokenizer.decode([ids[0]]))

    def on_tokens(toks):
        pieces.extend(int(t) for t in toks)
        if stream:
            stream(tokenizer.decode(toks))

    predict_ids(engine, ids, max_length, eos_id=eos_id, on_tokens=on_tokens, **kw)
    return tokenizer.decode(pieces)
