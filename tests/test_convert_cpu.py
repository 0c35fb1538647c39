"""Checkpoint converters (simplellminference_b200/convert.py): llama2.c .bin and HF safetensors -> the reference's raw fp32 blob.

The check that matters: a llama2.c checkpoint rotates INTERLEAVED pairs, the reference rotates (j, j + hd/2). After the converter's
row permutation of wq / wk, the reference CPU path (the oracle) must give the logits of a plain numpy forward with llama2.c's
interleaved RoPE on the ORIGINAL tensors (with the reference's own sigmoid(gate)*up, which the converter cannot change)."""
import json
import struct

import numpy as np
import pytest

from conftest import oracle_shape
from simplellminference_b200 import convert


def _tiny_llama2c(path, seed=0, shared=True):
    rng = np.random.default_rng(seed)
    dim, hidden, L, H, KVH, V, S = 64, 160, 2, 4, 2, 96, 24
    hd = dim // H
    kv = KVH * hd
    r = lambda *shp, s=1.0: (rng.standard_normal(shp) * s).astype(np.float32)   # noqa: E731
    t = dict(emb=r(V, dim), att_norm=1 + r(L, dim, s=0.05), wq=r(L, dim, dim, s=dim ** -0.5), wk=r(L, kv, dim, s=dim ** -0.5),
             wv=r(L, kv, dim, s=dim ** -0.5), wo=r(L, dim, dim, s=dim ** -0.5), ffn_norm=1 + r(L, dim, s=0.05),
             w1=r(L, hidden, dim, s=dim ** -0.5), w2=r(L, dim, hidden, s=hidden ** -0.5), w3=r(L, hidden, dim, s=dim ** -0.5),
             final_norm=1 + r(dim, s=0.05))
    with open(path, "wb") as f:
        f.write(struct.pack("<7i", dim, hidden, L, H, KVH, V if shared else -V, S))
        for k in ("emb", "att_norm", "wq", "wk", "wv", "wo", "ffn_norm", "w1", "w2", "w3", "final_norm"):
            f.write(t[k].tobytes())
        f.write(np.zeros(S * hd, np.float32).tobytes())   # freq_cis_real + freq_cis_imag
        if not shared:
            f.write(r(V, dim).tobytes())
    return t, dict(dim=dim, hidden=hidden, L=L, H=H, KVH=KVH, V=V, S=S, hd=hd)


def _llama2c_forward(t, c, tokens):
    """float64 forward with llama2.c's interleaved RoPE and the reference's sigmoid gate; returns the logits of every position."""
    f8 = {k: v.astype(np.float64) for k, v in t.items()}
    hd, H, KVH, L = c["hd"], c["H"], c["KVH"], c["L"]
    g = H // KVH
    kc = np.zeros((L, len(tokens), KVH * hd))
    vc = np.zeros_like(kc)
    out = []

    def norm(x, w):
        return x / np.sqrt((x * x).mean() + 1e-5) * w

    def rope(vec, pos, heads):
        vec = vec.reshape(heads, hd // 2, 2).copy()
        freq = 1.0 / (10000.0 ** (2.0 * np.arange(hd // 2) / hd))
        cs, sn = np.cos(pos * freq), np.sin(pos * freq)
        a, b = vec[:, :, 0].copy(), vec[:, :, 1].copy()
        vec[:, :, 0], vec[:, :, 1] = a * cs - b * sn, a * sn + b * cs
        return vec.reshape(-1)

    for pos, tok in enumerate(tokens):
        x = f8["emb"][tok].copy()
        for l in range(L):
            xn = norm(x, f8["att_norm"][l])
            q = rope(f8["wq"][l] @ xn, pos, H)
            kc[l, pos] = rope(f8["wk"][l] @ xn, pos, KVH)
            vc[l, pos] = f8["wv"][l] @ xn
            att = np.zeros(H * hd)
            for h in range(H):
                kh = kc[l, :pos + 1, (h // g) * hd:(h // g + 1) * hd]
                s = kh @ q[h * hd:(h + 1) * hd] / np.sqrt(hd)
                p = np.exp(s - s.max())
                p /= p.sum()
                att[h * hd:(h + 1) * hd] = p @ vc[l, :pos + 1, (h // g) * hd:(h // g + 1) * hd]
            x = x + f8["wo"][l] @ att
            xn = norm(x, f8["ffn_norm"][l])
            x = x + f8["w2"][l] @ ((1.0 / (1.0 + np.exp(-(f8["w1"][l] @ xn)))) * (f8["w3"][l] @ xn))
        out.append(f8["emb"] @ norm(x, f8["final_norm"]))
    return np.array(out)


def test_llama2c_checkpoint_through_the_reference_path(tmp_path, port):
    path = str(tmp_path / "tiny.bin")
    t, c = _tiny_llama2c(path)
    shape, blob, notes = convert.llama2c_to_blob(path)
    assert notes == [] and blob.dtype == np.float32
    assert (shape.vocab, shape.head_dim, shape.hidden, shape.kv_hidden, shape.inter, shape.max_len, shape.layers, shape.heads, shape.kv_heads) == \
        (c["V"], c["hd"], c["dim"], c["KVH"] * c["hd"], c["hidden"], c["S"], c["L"], c["H"], c["KVH"])
    assert blob.size == port.blob_floats(oracle_shape(shape))
    tokens = [1, 17, 5, 80, 33, 2, 64, 9]
    want = _llama2c_forward(t, c, tokens)
    om = port.model(oracle_shape(shape), blob)
    for pos, tok in enumerate(tokens):
        got = om.forward(tok, pos)
        assert float(np.abs(got - want[pos]).max()) <= 2e-4 * max(1.0, float(np.abs(want[pos]).max())), pos
    # without the permutation the scores differ: the check above is not vacuous
    raw = dict(emb=t["emb"], att_norm=t["att_norm"], ffn_norm=t["ffn_norm"], final_norm=t["final_norm"], wq=t["wq"], wk=t["wk"], wv=t["wv"],
               wo=t["wo"], up=t["w3"], gate=t["w1"], down=t["w2"])
    om2 = port.model(oracle_shape(shape), convert.assemble_blob(shape, raw))
    for pos, tok in enumerate(tokens):
        bad = om2.forward(tok, pos)
    assert float(np.abs(bad - want[-1]).max()) > 1e-2


def test_llama2c_untied_classifier_is_reported(tmp_path):
    path = str(tmp_path / "untied.bin")
    _tiny_llama2c(path, shared=False)
    shape, blob, notes = convert.llama2c_to_blob(path)
    assert len(notes) == 1 and "untied" in notes[0]


def _write_safetensors(path, tensors):
    header, blobs, off = {}, [], 0
    for name, (dt, arr) in tensors.items():
        if dt == "BF16":
            raw = (np.ascontiguousarray(arr, np.float32).view(np.uint32) >> 16).astype("<u2").tobytes()
        else:
            raw = np.ascontiguousarray(arr, dtype={"F32": "<f4", "F16": "<f2"}[dt]).tobytes()
        header[name] = {"dtype": dt, "shape": list(arr.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    h = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for b in blobs:
            f.write(b)


def test_hf_safetensors_to_blob(tmp_path, port):
    rng = np.random.default_rng(1)
    cfg = dict(hidden_size=32, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=48, vocab_size=40,
               max_position_embeddings=16, rms_norm_eps=1e-5, rope_theta=10000.0)
    d, L, kv, I, V = 32, 2, 16, 48, 40
    bf = lambda a: ((a.astype(np.float32).view(np.uint32) >> 16) << 16).view(np.float32)   # noqa: E731  values exactly representable in bf16
    tensors = {"model.embed_tokens.weight": ("BF16", bf(rng.standard_normal((V, d)))), "model.norm.weight": ("F32", 1 + 0.1 * rng.standard_normal(d)),
               "lm_head.weight": ("F16", rng.standard_normal((V, d)).astype(np.float16).astype(np.float32))}
    for l in range(L):
        p = f"model.layers.{l}."
        tensors[p + "input_layernorm.weight"] = ("F32", 1 + 0.1 * rng.standard_normal(d))
        tensors[p + "post_attention_layernorm.weight"] = ("F32", 1 + 0.1 * rng.standard_normal(d))
        for nm, shp in (("self_attn.q_proj", (d, d)), ("self_attn.k_proj", (kv, d)), ("self_attn.v_proj", (kv, d)), ("self_attn.o_proj", (d, d)),
                        ("mlp.up_proj", (I, d)), ("mlp.gate_proj", (I, d)), ("mlp.down_proj", (d, I))):
            tensors[p + nm + ".weight"] = ("BF16", bf(rng.standard_normal(shp) * 0.2))
    path = str(tmp_path / "model.safetensors")
    _write_safetensors(path, tensors)
    loaded = convert.read_safetensors(path)
    for name, (_, arr) in tensors.items():
        assert np.array_equal(loaded[name], np.asarray(arr, np.float32)), name
    shape, blob, notes = convert.hf_to_blob(loaded, cfg)
    assert len(notes) == 1 and "lm_head" in notes[0]
    sh = oracle_shape(shape)
    assert blob.size == port.blob_floats(sh)
    # segment by segment against the reference's tensor order (oracle/synth_weights.c: segments 0..8)
    off, cnt = port.segment(sh, 0)[:2]
    assert np.array_equal(blob[off:off + cnt].reshape(V, d), loaded["model.embed_tokens.weight"])
    off, cnt = port.segment(sh, 1)[:2]
    norms = blob[off:off + cnt].reshape(2 * L + 1, d)
    assert np.array_equal(norms[2], loaded["model.layers.1.input_layernorm.weight"]) and np.array_equal(norms[3], loaded["model.layers.1.post_attention_layernorm.weight"])
    assert np.array_equal(norms[2 * L], loaded["model.norm.weight"])
    for seg, nm, shp in ((2, "self_attn.q_proj", (d, d)), (3, "self_attn.k_proj", (kv, d)), (4, "self_attn.v_proj", (kv, d)), (5, "self_attn.o_proj", (d, d)),
                         (6, "mlp.up_proj", (I, d)), (7, "mlp.gate_proj", (I, d)), (8, "mlp.down_proj", (d, I))):
        off, cnt = port.segment(sh, seg)[:2]
        got = blob[off:off + cnt].reshape((L,) + shp)
        assert np.array_equal(got[1], loaded[f"model.layers.1.{nm}.weight"]), nm
    assert np.isfinite(port.model(sh, blob).forward(3, 0)).all()


def test_cli_writes_blob_and_sidecar(tmp_path):
    path = str(tmp_path / "tiny.bin")
    _tiny_llama2c(path)
    out = str(tmp_path / "tiny.f32")
    convert.main(["--llama2c", path, "--out", out])
    side = json.load(open(out + ".json"))
    blob = np.fromfile(out, dtype="<f4")
    assert blob.size == side["floats"] and side["hidden"] == 64 and side["layers"] == 2
