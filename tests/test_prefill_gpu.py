"""GPU parity tests of the batched prefill path (tcgen05/TMEM GEMMs + causal block attention).

The reference has no batched prefill: it runs its single-token forward once per prompt token (model.cpp:157-166).
That loop — the CPU oracle's — defines the expected KV-cache contents, last-position logits and next token.

Tolerances (floating point; the tensor-core operands are bf16, accumulation fp32):
* the GEMM alone against torch fp32 on the SAME bf16 operands: only the summation order differs -> 1e-4 * max|C|;
* attention alone against a torch fp32 softmax on the same bf16 q/K/V: P and the output are rounded to bf16
  (2^-9 relative each) -> 2e-2 * max|out|;
* whole prefill against the oracle: activations entering every GEMM are rounded to bf16 (2^-9 relative per operand).
  On the gain-1 synthetic model (see _blob): KV rows of every layer <= 1.2e-2 * max|row| (measured 5e-3), last-position
  logits <= 3e-3 * max|logit| (measured 2-5e-4), next token identical whenever the oracle's
  top1-top2 margin exceeds twice the logit error. On the repo-wide gain-4 model the same rounding is
  amplified chaotically layer by layer: layer 0 is held to 1e-2, deeper layers and logits are printed, not bounded;
* prefill against this engine's own token-by-token decode (fp32 activations): same bounds, and the decode that follows a
  prefill must keep running (state, position, history) exactly like the one that follows a token-by-token prompt.
"""
import os

import numpy as np
import pytest
import torch

from conftest import oracle_shape
from simplellminference_b200 import kernels as K
from simplellminference_b200.config import BF16, F32, ModelShape
from simplellminference_b200.engine import Engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bn", [128, 256, 0, 96, 224])
@pytest.mark.parametrize("T,N,K_", [(128, 128, 64), (128, 256, 256), (512, 1024, 4096), (300, 1000, 1408), (77, 264, 520), (1, 128, 64),
                                    (640, 4096 + 32, 2048)])
def test_gemm_tcgen05_matches_torch(T, N, K_, bn):
    g = torch.Generator(device="cuda").manual_seed(T * 131 + N * 7 + K_)
    a = torch.randn(T, K_, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K_, device="cuda", generator=g) / np.sqrt(K_)).to(torch.bfloat16)
    got = K.prefill_gemm(a, w, bn)
    torch.cuda.synchronize()
    want = a.float() @ w.float().T
    err = float((got - want).abs().max())
    assert err <= 1e-4 * max(1.0, float(want.abs().max())), (err, float(want.abs().max()))


def _attn_ref(q, kc, vc, pos0, heads, kv_heads):
    T = q.shape[0]
    hd = kc.shape[2]
    G = heads // kv_heads
    qf = q.float().view(T, heads, hd)
    out = torch.empty(T, heads, hd, device=q.device)
    n = pos0 + T
    mask = torch.arange(n, device=q.device)[None, :] <= (pos0 + torch.arange(T, device=q.device))[:, None]
    for h in range(heads):
        k = kc[h // G, :n].float().to(torch.bfloat16).float()   # the kernel multiplies bf16 K/V (an fp32 cache is rounded on load)
        v = vc[h // G, :n].float().to(torch.bfloat16).float()
        s = (qf[:, h] @ k.T) / np.sqrt(hd)
        s = s.masked_fill(~mask, float("-inf"))
        out[:, h] = torch.softmax(s, dim=-1) @ v
    return out.view(T, heads * hd)


@pytest.mark.parametrize("kv_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,pos0,hd,heads,kv_heads", [(64, 0, 128, 2, 2), (200, 0, 64, 4, 2), (129, 37, 128, 8, 2), (1, 5, 64, 2, 1), (511, 0, 128, 4, 4)])
def test_attention_matches_torch(T, pos0, hd, heads, kv_heads, kv_dtype):
    g = torch.Generator(device="cuda").manual_seed(T + 17 * pos0 + hd)
    S = pos0 + T + 3
    q = torch.randn(T, heads * hd, device="cuda", generator=g).to(torch.bfloat16)
    kc = torch.randn(kv_heads, S, hd, device="cuda", generator=g).to(kv_dtype)
    vc = torch.randn(kv_heads, S, hd, device="cuda", generator=g).to(kv_dtype)
    kc[:, pos0 + T:] = float("nan")   # rows past the block must never contribute
    vc[:, pos0 + T:] = float("nan")
    got = K.prefill_attention(q, kc, vc, pos0, heads, kv_heads).float()
    torch.cuda.synchronize()
    want = _attn_ref(q, kc, vc, pos0, heads, kv_heads)
    assert torch.isfinite(got).all()
    err = float((got - want).abs().max())
    assert err <= 2e-2 * float(want.abs().max()), err


SHAPES = {
    # name: (shape, kv dtype, prompt length)
    "gqa_hd64": (ModelShape(2048, 64, 512, 128, 1408, 320, 3, 8, 2), BF16, 200),
    "mha_hd128_bf16kv": (ModelShape(4096, 128, 1024, 1024, 2816, 400, 4, 8, 8), BF16, 300),
    "gqa_hd64_f32kv": (ModelShape(2048, 64, 512, 256, 1024, 160, 2, 8, 4), F32, 130),
}


def _prompt(n, vocab, seed):
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, vocab, size=n, dtype=np.int32)
    ids[0] = 1
    return ids


def _blob(port, ms, seed, gain):
    """Synthetic blob with bf16-exact values; gain 4 = the repo-wide synthetic model (projections std 4/sqrt(fan_in)), gain 1 =
    the same values with every projection matrix scaled by 0.25 (exact in bf16). The gain-4 model is deliberately
    chaotic (scores of std ~16 make the softmax amplify any perturbation layer by layer), which is right for token-identity
    tests of an fp32 path but turns the bf16 operand rounding of the tensor-core path into O(1e-1) differences after a
    few layers; the gain-1 blob measures the implementation instead of the chaos."""
    sh = oracle_shape(ms)
    blob = port.fill_blob(sh, seed, BF16)
    if gain != 4:
        for seg in range(2, 9):   # wq wk wv wo up gate down (model.cpp:372-462)
            off, cnt = port.segment(sh, seg)[:2]
            blob[off:off + cnt] *= gain / 4.0
    return blob


@pytest.mark.parametrize("gain", [1, 4])
@pytest.mark.parametrize("name", list(SHAPES))
def test_prefill_matches_oracle_and_decode_path(port, name, gain):
    ms, kvd, n = SHAPES[name]
    ids = _prompt(n, ms.vocab, 5)
    blob = _blob(port, ms, 21, gain)
    om = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1, kv_bf16=(kvd == BF16))
    for p in range(n - 1):
        om.step(int(ids[p]), p)
    want_l = om.forward(int(ids[n - 1]), n - 1)

    eng = Engine(ms, w_dtype=BF16, kv_dtype=kvd, mega=True).load_blob(blob)
    assert eng.mode == "megakernel" and eng.prefill_supported, eng.mode
    eng.prefill(ids)
    torch.cuda.synchronize()
    got_l = eng.buffer("model_pred").cpu().numpy()

    # the same prompt token by token through the decode step (fp32 activations) on a second engine
    dec = Engine(ms, w_dtype=BF16, kv_dtype=kvd, mega=True).load_blob(blob)
    dec.greedy(ids, n + 1)
    dec_l = dec.buffer("model_pred").cpu().numpy()

    kv = ms.kv_hidden
    worst = {}
    for l in range(ms.layers):
        for which, bid in (("k", 2), ("v", 3)):
            for p in (0, 1, n // 2, n - 2, n - 1):
                want = om.read(bid, (l * ms.max_len + p) * kv, kv)
                got = eng.kv_row(which, l, p).float().cpu().numpy()
                e = float(np.abs(got - want).max()) / max(1e-6, float(np.abs(want).max()))
                worst[l] = max(worst.get(l, 0.0), e)
    scale = max(1.0, float(np.abs(want_l).max()))
    err, err_dec = float(np.abs(got_l - want_l).max()), float(np.abs(got_l - dec_l).max())
    srt = np.sort(want_l)
    margin = float(srt[-1] - srt[-2])
    print(f"\n{name} gain {gain}: prefill vs oracle max|dlogit|={err:.3e} ({err / scale:.2e} of max|logit|), vs own decode {err_dec:.3e}; "
          f"kv rel err per layer { {k: round(v, 4) for k, v in worst.items()} }; oracle margin {margin:.3f}; "
          f"next token prefill {int(np.argmax(got_l))} oracle {int(np.argmax(want_l))}")
    assert worst[0] <= 1e-2, worst                      # layer 0: only the RMSNorm output was rounded to bf16
    if gain == 1:
        assert max(worst.values()) <= 1.2e-2, worst       # measured 5e-3, flat over the layers (one bf16 rounding of the GEMM operands)
        assert err <= 3e-3 * scale and err_dec <= 3e-3 * scale, (err, err_dec, scale)   # measured 2-5e-4 of max|logit| (the old 3e-2 was 60x looser)
        if margin > 2 * err:
            assert int(np.argmax(got_l)) == int(np.argmax(want_l))
    else:
        assert max(worst.values()) <= 0.5 and np.isfinite(got_l).all(), worst   # chaotic model: sanity only, numbers are printed

    # state after prefill == state after a token-by-token prompt: the decode that follows just keeps going
    eng.enqueue_steps(6)
    toks = eng.read_tokens(7)
    assert toks[0] == int(np.argmax(got_l))
    nxt_hist = eng.read_tokens(n + 6)
    assert np.array_equal(nxt_hist[:n - 1], ids[1:])       # history of the prompt part
    eng.close(); dec.close()


def test_prefill_in_two_blocks_equals_one(port):
    """prefill(a + b) == prefill(a) then prefill(b, start_pos=len(a)): the second block attends to the cached first one."""
    ms, kvd, n = SHAPES["gqa_hd64"]
    ids = _prompt(n, ms.vocab, 9)
    blob = _blob(port, ms, 4, 1)
    one = Engine(ms, w_dtype=BF16, kv_dtype=kvd, mega=True).load_blob(blob)
    two = Engine(ms, w_dtype=BF16, kv_dtype=kvd, mega=True).load_blob(blob)
    one.prefill(ids)
    two.prefill(ids[:77])
    two.prefill(ids[77:], start_pos=77)
    torch.cuda.synchronize()
    a, b = one.buffer("model_pred").cpu().numpy(), two.buffer("model_pred").cpu().numpy()
    # different block boundaries = different tile shapes / summation order in attention: tiny difference allowed
    assert float(np.abs(a - b).max()) <= 2e-2 * max(1.0, float(np.abs(a).max()))
    one.close(); two.close()


def test_prefill_rejected_where_unsupported():
    ms = ModelShape(512, 32, 128, 64, 384, 48, 3, 4, 2)
    eng = Engine(ms, w_dtype=F32, kv_dtype=F32, mega=True).load_synthetic(1)
    assert not eng.prefill_supported
    with pytest.raises(Exception):
        eng.prefill([1, 2, 3])
    assert np.array_equal(eng.greedy([1, 2, 3], 5)[:2], [2, 3])   # the token-by-token path still serves the prompt
    eng.close()


def test_prefill_longer_than_one_block(port):
    """Prompts longer than the 1024-row block of the workspace run as several blocks (the later ones attend to the cached rows
    of the earlier ones); n = 1 is a degenerate block. Compared with the oracle on the gain-1 blob."""
    ms = ModelShape(1024, 64, 256, 128, 512, 1300, 2, 4, 2)
    n = 1100
    ids = _prompt(n, ms.vocab, 3)
    blob = _blob(port, ms, 8, 1)
    om = port.model(oracle_shape(ms), blob, threads=os.cpu_count() or 1, kv_bf16=True)
    first = om.forward(int(ids[0]), 0)
    for p in range(1, n - 1):
        om.step(int(ids[p]), p)
    want_l = om.forward(int(ids[n - 1]), n - 1)
    eng = Engine(ms, w_dtype=BF16, kv_dtype=BF16, mega=True).load_blob(blob)
    eng.prefill(ids[:1])
    torch.cuda.synchronize()
    got1 = eng.buffer("model_pred").cpu().numpy()
    assert float(np.abs(got1 - first).max()) <= 3e-2 * max(1.0, float(np.abs(first).max()))
    eng.prefill(ids)
    torch.cuda.synchronize()
    got = eng.buffer("model_pred").cpu().numpy()
    err = float(np.abs(got - want_l).max())
    assert err <= 3e-2 * max(1.0, float(np.abs(want_l).max())), err
    assert int(np.argmax(got)) == int(np.argmax(want_l))
    with pytest.raises(Exception):
        eng.prefill(ids, start_pos=ms.max_len - 10)      # would overrun the KV cache
    with pytest.raises(Exception):
        eng.prefill([ms.vocab])                           # token outside the vocabulary
    eng.close()
