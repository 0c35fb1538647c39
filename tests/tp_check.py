"""tests/tp_check.py — run under torchrun on N GPUs (tests/test_tp_gpu.py launches it): tensor-parallel engines
(one rank per GPU, NCCL) must produce the oracle's greedy token stream and logits on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from simplellminference_b200.config import PRESETS, ModelShape, F32, BF16  # noqa: E402
from simplellminference_b200.engine import Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    port = loader.Port()
    os.chdir("/tmp")
    cases = [("tiny_gqa", PRESETS["tiny_gqa"], F32, F32, [1, 7, 300], 40),
             ("medium_gqa_bf16", ModelShape(4096, 128, 1024, 512, 2816, 160, 4, 8, 4), BF16, BF16, list(range(1, 17)), 120)]
    for name, ms, wd, kvd, prompt, n_total in cases:
        if ms.kv_heads % world or ms.heads % world:
            continue
        shape = loader.Shape(ms.vocab, ms.head_dim, ms.hidden, ms.kv_hidden, ms.inter, ms.max_len, ms.layers, ms.heads, ms.kv_heads, ms.eps, ms.theta)
        blob = port.fill_blob(shape, 1234, wd, 64)
        want, want_l = port.model(shape, blob, threads=4, kv_bf16=(kvd == BF16)).greedy(prompt, n_total)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        eng = Engine(ms, w_dtype=wd, kv_dtype=kvd, tp_rank=rank, tp_size=world, stream=stream).load_synthetic(1234).init_comm(dist)
        got = eng.greedy(prompt, n_total)
        assert np.array_equal(got, want), (name, rank, got[:10], want[:10])
        v_loc = ms.vocab // world
        logits = eng.buffer("model_pred").cpu().numpy()
        err = float(np.abs(logits - want_l[rank * v_loc:(rank + 1) * v_loc]).max())
        tol = (5e-3 if kvd == BF16 else 3e-4) * max(1.0, float(np.abs(want_l).max()))
        assert err <= tol, (name, rank, err, tol)
        eng.close()
        if rank == 0:
            print(f"tp{world} {name}: {n_total - 1} tokens identical on every rank, max|dlogit|={err:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("TP_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
