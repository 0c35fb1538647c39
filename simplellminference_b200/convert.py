"""Checkpoint converters into the reference's weight file.

The reference mmaps a HEADERLESS raw little-endian fp32 blob in its own tensor order (source/model/model.cpp:340-468):
    E[V][d] | (2L+1) norm vectors [d] (attention norm l, ffn norm l, ..., final) | wq[L][d][d] | wk[L][kv][d] | wv[L][kv][d] |
    wo[L][d][d] | up[L][I][d] | gate[L][I][d] | down[L][d][I]
with the classifier TIED to E (model.cpp:350-352) and rotate-half RoPE (pairs (j, j+hd/2), rope_kernel.cpp:34-37). Its model
shape is hard-coded (include/model/config.h:5-17), so every converter here also returns / writes the shape as a JSON side-car.

  llama2.c `.bin` (the format stories15M / stories110M ship in): 7-int32 header (dim, hidden_dim, n_layers, n_heads, n_kv_heads,
      +-vocab_size, seq_len; negative vocab = untied classifier), then fp32 tensors in llama2.c's order. llama2.c rotates INTERLEAVED
      pairs (2i, 2i+1): the rows of every wq / wk head are permuted to [0, 2, 4, ..., 1, 3, 5, ...] so that rotate-half on the
      permuted q, k gives the same attention scores (q.k is invariant under a common permutation) — the same permutation HF's
      conversion script applies.
  HF safetensors (Llama family, already rotate-half): tensors looked up by their HF names, f32 / f16 / bf16 → fp32. Parsed here
      (8-byte header length + JSON + raw data), no dependency on the `safetensors` package.

An untied classifier (lm_head != embed_tokens) cannot be expressed in the reference's format: it is reported and dropped.
Note: the reference multiplies up by sigmoid(gate), not silu(gate) (swiglu_kernel.cpp:12-13), so a real checkpoint converted
here reproduces the REFERENCE's outputs for that checkpoint, not the original model's.

CLI:  python -m simplellminference_b200.convert --llama2c stories15M.bin --out stories15M.f32 [--config-out stories15M.json]
      python -m simplellminference_b200.convert --safetensors model.safetensors --hf-config config.json --out llama.f32
"""
from __future__ import annotations

import argparse
import json
import struct

import numpy as np

from .config import ModelShape


def _rope_permute(w: np.ndarray, n_heads: int) -> np.ndarray:
    """Rows of each head from interleaved-pair order to rotate-half order: [0, 2, 4, ..., 1, 3, 5, ...]."""
    rows, cols = w.shape
    hd = rows // n_heads
    return w.reshape(n_heads, hd // 2, 2, cols).transpose(0, 2, 1, 3).reshape(rows, cols)


def assemble_blob(shape: ModelShape, t: dict) -> np.ndarray:
    """t: emb [V][d], att_norm [L][d], ffn_norm [L][d], final_norm [d], wq [L][d][d], wk/wv [L][kv][d], wo [L][d][d],
    up/gate [L][I][d], down [L][d][I] (already in the reference's conventions) -> the reference's flat fp32 blob."""
    L, d, kv, I, V = shape.layers, shape.hidden, shape.kv_hidden, shape.inter, shape.vocab
    want = dict(emb=(V, d), att_norm=(L, d), ffn_norm=(L, d), final_norm=(d,), wq=(L, d, d), wk=(L, kv, d), wv=(L, kv, d), wo=(L, d, d),
                up=(L, I, d), gate=(L, I, d), down=(L, d, I))
    for k, shp in want.items():
        if tuple(t[k].shape) != shp:
            raise ValueError(f"{k}: shape {tuple(t[k].shape)}, expected {shp}")
    norms = np.empty((2 * L + 1, d), np.float32)
    norms[0:2 * L:2] = t["att_norm"]
    norms[1:2 * L:2] = t["ffn_norm"]
    norms[2 * L] = t["final_norm"]
    parts = [t["emb"], norms, t["wq"], t["wk"], t["wv"], t["wo"], t["up"], t["gate"], t["down"]]
    return np.concatenate([np.ascontiguousarray(p, dtype=np.float32).reshape(-1) for p in parts])


def read_llama2c(path: str):
    """-> (ModelShape, tensors dict in llama2.c conventions, shared_classifier flag)."""
    with open(path, "rb") as f:
        dim, hidden, n_layers, n_heads, n_kv_heads, vocab, seq_len = struct.unpack("<7i", f.read(28))
        shared = vocab > 0
        vocab = abs(vocab)
        hd = dim // n_heads
        kv = n_kv_heads * hd

        def take(*shp):
            n = int(np.prod(shp))
            a = np.frombuffer(f.read(4 * n), dtype="<f4")
            if a.size != n:
                raise ValueError(f"{path}: truncated (wanted {n} floats for a {shp} tensor)")
            return a.reshape(shp)

        t = dict(emb=take(vocab, dim), att_norm=take(n_layers, dim), wq=take(n_layers, dim, dim), wk=take(n_layers, kv, dim),
                 wv=take(n_layers, kv, dim), wo=take(n_layers, dim, dim), ffn_norm=take(n_layers, dim), w1=take(n_layers, hidden, dim),
                 w2=take(n_layers, dim, hidden), w3=take(n_layers, hidden, dim), final_norm=take(dim))
        f.read(4 * seq_len * hd)   # freq_cis_real + freq_cis_imag (legacy; recomputed by everybody)
        if not shared:
            t["lm_head"] = take(vocab, dim)
    shape = ModelShape(vocab, hd, dim, kv, hidden, seq_len, n_layers, n_heads, n_kv_heads, eps=1e-5, theta=10000.0)
    return shape, t, shared


def llama2c_to_blob(path: str):
    """llama2.c checkpoint -> (ModelShape, reference blob, notes)."""
    shape, t, shared = read_llama2c(path)
    notes = []
    if not shared:
        notes.append("untied classifier (lm_head) dropped: the reference ties the classifier to the embedding table")
    ref = dict(emb=t["emb"], att_norm=t["att_norm"], ffn_norm=t["ffn_norm"], final_norm=t["final_norm"],
               wq=np.stack([_rope_permute(w, shape.heads) for w in t["wq"]]), wk=np.stack([_rope_permute(w, shape.kv_heads) for w in t["wk"]]),
               wv=t["wv"], wo=t["wo"], up=t["w3"], gate=t["w1"], down=t["w2"])   # llama2.c: w2(silu(w1 x) * w3 x)
    return shape, assemble_blob(shape, ref), notes


_ST_DTYPES = {"F32": ("<f4", 4), "F16": ("<f2", 2), "BF16": (None, 2)}


def read_safetensors(path: str) -> dict:
    """Minimal safetensors reader: name -> fp32 numpy array (F32 / F16 / BF16 tensors)."""
    out = {}
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n))
        base = 8 + n
        for name, info in header.items():
            if name == "__metadata__":
                continue
            dt, (b0, b1) = info["dtype"], info["data_offsets"]
            if dt not in _ST_DTYPES:
                raise ValueError(f"{name}: dtype {dt} not supported (F32, F16, BF16)")
            f.seek(base + b0)
            raw = f.read(b1 - b0)
            if dt == "BF16":
                a = (np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16).view(np.float32)
            else:
                a = np.frombuffer(raw, dtype=_ST_DTYPES[dt][0]).astype(np.float32)
            out[name] = a.reshape(info["shape"])
    return out


def hf_to_blob(tensors: dict, cfg: dict):
    """HF Llama tensors (already rotate-half) + config.json dict -> (ModelShape, reference blob, notes)."""
    d, L, H = cfg["hidden_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"]
    KVH = cfg.get("num_key_value_heads", H)
    hd = cfg.get("head_dim", d // H)
    shape = ModelShape(cfg["vocab_size"], hd, d, KVH * hd, cfg["intermediate_size"], cfg.get("max_position_embeddings", 2048), L, H, KVH,
                       eps=float(cfg.get("rms_norm_eps", 1e-5)), theta=float(cfg.get("rope_theta", 10000.0)))
    if H * hd != d:
        raise ValueError("the reference assumes heads * head_dim == hidden (wq and wo are d x d, model.cpp:372-420)")
    g = lambda name: tensors[name]   # noqa: E731
    lay = lambda fmt: np.stack([g(fmt.format(l)) for l in range(L)])   # noqa: E731
    ref = dict(emb=g("model.embed_tokens.weight"), att_norm=lay("model.layers.{}.input_layernorm.weight"),
               ffn_norm=lay("model.layers.{}.post_attention_layernorm.weight"), final_norm=g("model.norm.weight"),
               wq=lay("model.layers.{}.self_attn.q_proj.weight"), wk=lay("model.layers.{}.self_attn.k_proj.weight"),
               wv=lay("model.layers.{}.self_attn.v_proj.weight"), wo=lay("model.layers.{}.self_attn.o_proj.weight"),
               up=lay("model.layers.{}.mlp.up_proj.weight"), gate=lay("model.layers.{}.mlp.gate_proj.weight"),
               down=lay("model.layers.{}.mlp.down_proj.weight"))
    notes = []
    if "lm_head.weight" in tensors and not np.array_equal(tensors["lm_head.weight"], ref["emb"]):
        notes.append("untied classifier (lm_head.weight) dropped: the reference ties the classifier to the embedding table")
    return shape, assemble_blob(shape, ref), notes


def write_blob(path: str, blob: np.ndarray) -> None:
    np.ascontiguousarray(blob, dtype="<f4").tofile(path)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--llama2c", help="llama2.c .bin checkpoint")
    src.add_argument("--safetensors", help="HF Llama .safetensors file (single shard)")
    ap.add_argument("--hf-config", help="HF config.json (with --safetensors)")
    ap.add_argument("--out", required=True, help="raw fp32 blob in the reference's tensor order")
    ap.add_argument("--config-out", help="JSON side-car with the model shape (default: <out>.json)")
    a = ap.parse_args(argv)
    if a.llama2c:
        shape, blob, notes = llama2c_to_blob(a.llama2c)
    else:
        if not a.hf_config:
            ap.error("--safetensors needs --hf-config")
        with open(a.hf_config) as f:
            shape, blob, notes = hf_to_blob(read_safetensors(a.safetensors), json.load(f))
    write_blob(a.out, blob)
    with open(a.config_out or a.out + ".json", "w") as f:
        json.dump(dict(shape.as_dict(), floats=int(blob.size), notes=notes), f, indent=1)
    print(json.dumps(dict(out=a.out, floats=int(blob.size), shape=shape.as_dict(), notes=notes)))


if __name__ == "__main__":
    main()
