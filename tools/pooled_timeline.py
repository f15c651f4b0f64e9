"""Forward timeline of the pooled cfg5 step on every rank (torchrun; MSCS_FWD_TIMELINE=1): milliseconds between the
launches of the forward sweeps as seen by CUDA events on each rank's stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import mscs_b200
from mscs_b200 import synth, _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = synth.CONFIGS["cfg5"]
labels_all = synth.make_labels(cfg)
n = cfg["n"]; nl = n // world
fts = []
for si, st in enumerate(cfg["strides"]):
    f = torch.empty((nl, cfg["C"], cfg["H"] // st, cfg["W"] // st), device=dev)
    for b in range(rank * nl, (rank + 1) * nl):
        g = torch.Generator(device=dev); g.manual_seed(100003 * (si + 1) + b)
        f[b - rank * nl].normal_(generator=g)
    fts.append(f.requires_grad_(True))
lab = labels_all[rank * nl:(rank + 1) * nl].contiguous().to(dev)
mod = mscs_b200.DenseContrastiveLossV2_ms(dict(cfg["loss"]), comm=mscs_b200.TorchDistComm())
torch.manual_seed(0)
def step():
    for f in fts: f.grad = None
    mod(lab, fts).backward()
for _ in range(3): step()
lib = _lib.load()
buf = np.zeros(8, np.float32); acc = np.zeros(8); K = 8
for _ in range(K):
    dist.barrier(); torch.cuda.synchronize()
    step()
    k = lib.mscs_debug_fwd_timeline(buf.ctypes.data, 8)
    acc[:k] += buf[:k]
# steady state (no synchronisation between steps, as in bench.py): stage times of THIS rank + the timeline of the last step
from mscs_b200 import _ops
dist.barrier(); torch.cuda.synchronize()
_ops.TIMING = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
stage = {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in _ops.TIMING.items()}
_ops.TIMING = None
k = lib.mscs_debug_fwd_timeline(buf.ctypes.data, 8)
print(f"\nrank {rank} steady: {e0.elapsed_time(e1) / 10:.3f} ms/step, stages {stage}, last step's sweeps (us) {[round(float(x) * 1e3, 1) for x in buf[:k]]}", flush=True)
spans = np.zeros((160, 4), np.uint64)
for mode in (0, 1):
    if lib.mscs_debug_cta_spans_fwd(spans.ctypes.data, mode) == 0:
        break
    sp = spans[:148].astype(np.int64)
    t0, t1, cyc, sm = sp[:, 0], sp[:, 1], sp[:, 2], sp[:, 3]
    dur = (t1 - t0) / 1e3
    order = np.argsort(-dur)[:3]
    print(f"\nrank {rank} sweep{mode} CTAs: kernel span {(t1.max() - t0.min()) / 1e3:.1f} us | start spread {(t0.max() - t0.min()) / 1e3:.1f} us | "
          f"duration min/median/max {dur.min():.1f}/{np.median(dur):.1f}/{dur.max():.1f} us | clock {np.median(cyc / (t1 - t0)):.3f} GHz | "
          f"slowest CTAs (cta, sm, us, start offset us): {[(int(i), int(sm[i]), round(float(dur[i]), 1), round(float(t0[i] - t0.min()) / 1e3, 1)) for i in order]} | "
          f"distinct SMs {len(set(sm.tolist()))}", flush=True)
print(f"\nrank {rank}: row_ranges, work tables, sweep0, sweep1 (us) = {[round(float(x) / K * 1e3, 1) for x in acc[:k]]}", flush=True)
dist.destroy_process_group()
