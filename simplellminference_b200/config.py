"""Model-shape presets (reference: include/model/config.h:5-17 hard-codes ONE shape; BASELINE.json names five)."""
from __future__ import annotations

from dataclasses import dataclass, asdict

F32, BF16, INT8 = 0, 1, 2
DTYPE_NAMES = {F32: "f32", BF16: "bf16", INT8: "int8"}


@dataclass(frozen=True)
class ModelShape:
    """Field-for-field the reference's model::LlamaModelConfig."""

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

    def __post_init__(self):
        if self.heads * self.head_dim != self.hidden:
            raise ValueError("heads*head_dim must equal hidden (the reference's wq/wo are d x d, model.cpp:372-378)")
        if self.kv_heads * self.head_dim != self.kv_hidden or self.heads % self.kv_heads:
            raise ValueError("inconsistent kv dims")

    def n_params(self) -> int:
        d, kv, I, L, V = self.hidden, self.kv_hidden, self.inter, self.layers, self.vocab
        return V * d + (2 * L + 1) * d + L * (2 * d * d + 2 * kv * d + 3 * I * d)

    def step_bytes(self, pos: int, w_dtype: int = BF16, kv_dtype: int = BF16, group: int = 64) -> float:
        """Algorithmic HBM bytes of one decode step at position `pos`."""
        bw = {F32: 4.0, BF16: 2.0, INT8: 1.0 + 4.0 / group}[w_dtype]
        bkv = 4.0 if kv_dtype == F32 else 2.0
        d, kv, I, L, V = self.hidden, self.kv_hidden, self.inter, self.layers, self.vocab
        return (bw * (V * d + L * (2 * d * d + 2 * kv * d + 3 * I * d)) + 4.0 * (2 * L + 1) * d + bw * d
                + bkv * 2 * L * kv * (pos + 1) + bkv * 2 * L * kv)

    def as_dict(self):
        return asdict(self)


# BASELINE.json configs (shapes only; weights are random-init)
PRESETS = {
    "stories15M": ModelShape(32000, 48, 288, 288, 768, 256, 6, 6, 6),
    "stories110M": ModelShape(32000, 64, 768, 768, 2048, 1024, 12, 12, 12),
    "tinyllama-1.1b": ModelShape(32000, 64, 2048, 256, 5632, 2048, 22, 32, 4),
    "llama2-7b": ModelShape(32000, 128, 4096, 4096, 11008, 4096, 32, 32, 32),
    # +8 positions of slack: the reference's GQA RoPE over-run limits the parity domain to pos <= S - H/KVH
    "llama3-8b": ModelShape(128256, 128, 4096, 1024, 14336, 8192 + 8, 32, 32, 8, theta=500000.0),
    # small shapes for tests
    "tiny_gqa": ModelShape(512, 32, 128, 64, 384, 48, 3, 4, 2),
    "tiny_mha_hd48": ModelShape(300, 48, 96, 96, 200, 40, 2, 2, 2),
}
