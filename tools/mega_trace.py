"""tools/mega_trace.py — per-phase time decomposition of the megakernel from its %globaltimer stamps."""
import argparse, dataclasses, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simplellminference_b200.config import PRESETS, BF16, F32
from simplellminference_b200.engine import Engine
ap = argparse.ArgumentParser(); ap.add_argument("--layers", type=int, default=4); ap.add_argument("--pos", type=int, default=512); ap.add_argument("--ll", action="store_true"); ap.add_argument("--fuse-down", action="store_true"); ap.add_argument("--v2", action="store_true"); ap.add_argument("--calibrate", action="store_true")
ap.add_argument("--config", default="llama2-7b"); ap.add_argument("--wdtype", default="bf16"); ap.add_argument("--kvdtype", default="bf16")
a = ap.parse_args()
ms = dataclasses.replace(PRESETS[a.config], layers=a.layers)
# under torchrun (WORLD_SIZE > 1): the tensor-parallel word-based kernel, one rank per GPU; rank 0 prints its own timeline
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
eng = Engine(ms, w_dtype={'bf16': BF16, 'f32': F32}[a.wdtype], kv_dtype={'bf16': BF16, 'f32': F32}[a.kvdtype], stream=stream, mega=True, mega_ll=a.ll, mega_fuse_down=a.fuse_down, mega_v2=a.v2,
             tp_rank=rank, tp_size=world, p2p_allreduce=(world > 1)).load_synthetic(1)
if world > 1:
    eng.init_p2p(dist)
    a.ll = True   # the only megakernel under tensor parallelism
if a.calibrate: eng.calibrate(3)
print("mode:", eng.mode, "calibrated" if a.calibrate else "")
eng.set_state(1, a.pos); eng.enqueue_steps(3); torch.cuda.synchronize()
tr = eng.buffer(200); torch.cuda.synchronize()
eng.enqueue_steps(1); torch.cuda.synchronize()
t = tr.view(torch.int64).cpu().numpy().reshape(-1, 512, 8)
nev = 5 * a.layers + 1
if a.ll:   # no barriers in the word-based kernel: 'barrier' = time from results sent to the start of this CTA's next phase
    for ev in range(nev - 1):
        t[:, ev, 5] = t[:, ev + 1, 0]
names = ["qkv", "att", "wo", "gate_up", "down"]
t0 = t[:, 0, 0].min()
print(f"grid={t.shape[0]} step total = {(t[:, nev-1, :5].max() - t0)/1e3:.1f} us")
agg = {}
for ev in range(nev):
    kind = "cls" if ev == nev - 1 else names[ev % 5]
    s = t[:, ev, :].astype(np.float64)
    start = s[:, 0]; 
    row = dict(begin=(start.min() - t0) / 1e3, skew=(start.max() - start.min()) / 1e3)
    if kind != "att":
        row.update(prologue=(s[:, 1] - s[:, 0]).mean() / 1e3, stream_mean=(s[:, 3] - s[:, 1]).mean() / 1e3, stream_max=(s[:, 3] - s[:, 1]).max() / 1e3,
                   stream_min=(s[:, 3] - s[:, 1]).min() / 1e3, epilogue=(s[:, 4] - s[:, 3]).mean() / 1e3)
    else:
        row.update(work_mean=(s[:, 4] - s[:, 0]).mean() / 1e3, work_max=(s[:, 4] - s[:, 0]).max() / 1e3)
    if kind != "cls":
        row.update(barrier_mean=(s[:, 5] - s[:, 4]).mean() / 1e3, phase_total=(s[:, 5].max() - start.min()) / 1e3)
    agg.setdefault(kind, []).append(row)
for kind, rows in agg.items():
    keys = [k for k in rows[0] if k != "begin"]
    print(kind.ljust(8), " ".join(f"{k}={np.mean([r[k] for r in rows[1:] or rows]):7.2f}" for k in keys))

# per-CTA streaming time of the big phases, layer by layer: is the spread systematic (same CTAs slow every time)?
smid = t[:, 0, 7]
import json as _json
per = {}
for kind, off in (("qkv", 0), ("gate_up", 3), ("down", 4)):
    m = np.stack([(t[:, 5 * l + off, 3] - t[:, 5 * l + off, 1]).astype(np.float64) / 1e3 for l in range(1, a.layers)], 1)   # [cta][layer]
    per[kind] = m
    mean = m.mean(1)
    cc = np.corrcoef(m.T)
    print(f"{kind}: per-CTA mean stream us min {mean.min():.2f} max {mean.max():.2f} std {mean.std():.2f}; layer-to-layer corr of per-CTA times: {cc[np.triu_indices_from(cc, 1)].mean():.2f}")
allm = np.concatenate([per[k] / per[k].mean() for k in per], 1).mean(1)
order = np.argsort(allm)
print("slowest CTAs (cta, smid, relative time):", [(int(c), int(smid[c]), round(float(allm[c]), 3)) for c in order[-10:]])
print("fastest CTAs:", [(int(c), int(smid[c]), round(float(allm[c]), 3)) for c in order[:10]])
print("relative time by smid parity / half:", {"smid<74": round(float(allm[smid < 74].mean()), 4), "smid>=74": round(float(allm[smid >= 74].mean()), 4),
      "even": round(float(allm[smid % 2 == 0].mean()), 4), "odd": round(float(allm[smid % 2 == 1].mean()), 4)})
np.save("gpurun_out/mega_rel_time.npy", np.stack([smid.astype(np.float64), allm], 1)) if (os.path.isdir("gpurun_out") and rank == 0) else None
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); eng.close(); dist.destroy_process_group()
