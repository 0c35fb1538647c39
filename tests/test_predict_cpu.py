"""The predict() driver and the tokenizer boundary (simplellminference_b200/predict.py) on the CPU: a real sentencepiece model
(third-party, trained here on a toy corpus) and a stand-in engine with the Engine call surface, so the control flow — prompt
echo, chunked generation, EOS stop, streaming — is checked without a GPU. The GPU test of the same driver is in
tests/test_engine_gpu.py::test_predict_driver."""
import os
import types

import numpy as np
import pytest

from simplellminference_b200.predict import SPELayer, predict, predict_ids

spm = pytest.importorskip("sentencepiece")


class ToyEngine:
    """next token = (7 * token + 3) mod V, with the Engine methods predict_ids uses (greedy / prefill / enqueue_steps / read_tokens)."""

    def __init__(self, vocab=64, max_len=128, prefill_supported=True):
        self.shape = types.SimpleNamespace(vocab=vocab, max_len=max_len)
        self.prefill_supported = prefill_supported
        self.history, self.token, self.calls = [], None, []

    def _next(self, t):
        return (7 * int(t) + 3) % self.shape.vocab

    def greedy(self, prompt, n_total):
        self.calls.append("greedy")
        self.history = [int(t) for t in prompt[1:]]
        self.token = int(prompt[-1])
        while len(self.history) < n_total - 1:
            self.token = self._next(self.token)
            self.history.append(self.token)
        return np.asarray(self.history, np.int32)

    def prefill(self, prompt):
        self.calls.append("prefill")
        self.history = [int(t) for t in prompt[1:]] + [self._next(prompt[-1])]
        self.token = self.history[-1]

    def enqueue_steps(self, k):
        self.calls.append(("steps", k))
        for _ in range(k):
            self.token = self._next(self.token)
            self.history.append(self.token)

    def read_tokens(self, k):
        return np.asarray(self.history[-k:], np.int32)


def _reference_stream(prompt, n, vocab=64):
    out, t = [int(x) for x in prompt[1:]], int(prompt[-1])
    while len(out) < n:
        t = (7 * t + 3) % vocab
        out.append(t)
    return out


@pytest.mark.parametrize("prefill,ask", [(True, True), (True, False), (False, True), (False, False)])
def test_predict_ids_matches_the_token_loop(prefill, ask):
    """The prompt goes token by token through the decode step (the reference's loop, the parity path) unless the caller opts in to the
    batched tensor-core pass AND the engine supports it."""
    eng = ToyEngine(prefill_supported=prefill)
    got = predict_ids(eng, [5, 9, 2], 40, chunk=7, batched_prefill=ask)
    assert got.tolist() == _reference_stream([5, 9, 2], 40)
    assert eng.calls[0] == ("prefill" if (prefill and ask) else "greedy")
    assert predict_ids(ToyEngine(prefill_supported=True), [5, 9, 2], 12).size == 12 and True
    assert sum(k for c in eng.calls[1:] for k in [c[1]]) == 40 - 3      # chunks add up exactly to max_length


def test_predict_ids_stops_at_eos_and_validates():
    full = _reference_stream([5], 60)
    eos = full[20]
    first = full.index(eos)
    got = predict_ids(ToyEngine(), [5], 60, eos_id=eos, chunk=8)
    assert got.tolist() == full[:first + 1] and got[-1] == eos
    assert predict_ids(ToyEngine(), [5], 60, eos_id=None, chunk=8).size == 60      # the reference never stops early
    with pytest.raises(ValueError):
        predict_ids(ToyEngine(max_len=32), [1], 32)                                 # max_length must be below the context
    with pytest.raises(ValueError):
        predict_ids(ToyEngine(), [], 8)
    with pytest.raises(ValueError):
        predict_ids(ToyEngine(), [1, 2, 3, 4], 3)


@pytest.fixture(scope="module")
def tokenizer(tmp_path_factory):
    d = tmp_path_factory.mktemp("spm")
    corpus = d / "corpus.txt"
    corpus.write_text("\n".join(["the quick brown fox jumps over the lazy dog", "a stitch in time saves nine", "to be or not to be that is the question",
                                 "all that glitters is not gold", "the early bird catches the worm"] * 20))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(d / "toy"), vocab_size=60, model_type="bpe", minloglevel=2)
    return SPELayer(str(d / "toy.model"))


def test_spelayer_surface(tokenizer, tmp_path):
    ids = tokenizer.encode("the quick brown fox")
    assert ids and all(0 <= i < tokenizer.GetVocabularySize() for i in ids)
    assert tokenizer.decode(ids) == "the quick brown fox"
    assert tokenizer.GetVocabularySize() == 60
    with pytest.raises(Exception):
        SPELayer(str(tmp_path / "missing.model"))                                  # encode.cpp:7-10: load failure is an error


def test_predict_text_streams_and_returns_text(tokenizer):
    eng = ToyEngine(vocab=64)
    chunks = []
    text = predict(eng, tokenizer, "the lazy dog", 24, stream=chunks.append, chunk=5)
    ids = tokenizer.encode("the lazy dog")
    want = [ids[0]] + _reference_stream(ids, 24)           # ids 60..63 exist in the model but not in the tokenizer: no text
    assert text == tokenizer.decode(want) and text.startswith("the lazy dog")
    assert "".join(chunks).replace(" ", "") == text.replace(" ", "")               # the streamed pieces are the same text
    assert len(chunks) >= 3                                                         # first piece, prompt echo + first token, then chunks
    with pytest.raises(ValueError):
        predict(ToyEngine(vocab=10), tokenizer, "the", 8)                           # tokenizer larger than the model's vocabulary
