#!/usr/bin/env python
"""tools/gemm_bench.py — the prefill GEMM kernel alone (row-major W) against cuBLAS (torch.matmul) on the llama2-7b prefill shapes.
cuBLAS is only the yardstick here; nothing in the product path calls it."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from simplellminference_b200 import _lib, kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=512)
ap.add_argument("--bn", type=int, default=0)
ap.add_argument("--pdl", type=int, default=1)
ap.add_argument("--pair", type=int, default=-1)
a = ap.parse_args()
_lib.load().sllm_tune(2, a.pair)
_lib.load().sllm_tune(3, a.pdl)
T = a.tokens
shapes = [("qkv", 12288, 4096), ("wo", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008)]
nw = 24   # rotate over several weight matrices so that W comes from HBM like in the real layer loop
for name, N, Kd in shapes:
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(T, Kd, device="cuda", generator=g).to(torch.bfloat16)
    Ws = [(torch.randn(N, Kd, device="cuda", generator=g) / Kd ** 0.5).to(torch.bfloat16) for _ in range(nw)]
    res = {"gemm": name, "T": T, "N": N, "K": Kd}
    for which in ("ours", "cublas"):
        fn = (lambda W: K.prefill_gemm(A, W, a.bn)) if which == "ours" else (lambda W: torch.matmul(A, W.T))
        for W in Ws[:3]:
            fn(W)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for W in Ws:
            fn(W)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / nw
        res[which + "_us"] = round(us, 1)
        res[which + "_tflops"] = round(2.0 * T * N * Kd / us / 1e6, 1)
    print(json.dumps(res))
