"""Event trace of one CTA of sweep 0 of the forward tensor kernel (library built with `make trace`, MSCS_LIB set)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mscs_b200
from mscs_b200 import synth, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["cfg2"]
labels, feats = synth.make_inputs("cfg2")
mod = mscs_b200.DenseContrastiveLossV2_ms(dict(cfg["loss"]))
labels = labels.to(dev); fg = [f.to(dev).requires_grad_(True) for f in feats]
torch.manual_seed(0)
def step():
    for f in fg: f.grad = None
    mod(labels, fg).backward()
for _ in range(3): step()
buf = np.zeros(8192, np.uint64)
lib.mscs_debug_trace_fwd(buf.ctypes.data, 8192)
step()
lib.mscs_debug_trace_fwd(buf.ctypes.data, 8192)
tr = buf.astype(np.int64).reshape(4, 256, 8)
mma, e4, e11, e12 = tr[0], tr[1], tr[2], tr[3]
n = int((mma[:, 2] > 0).sum())
print("tiles traced:", n)
m = lambda x: int(np.mean(x))
lo, hi = 10, n - 4
T = np.arange(lo, hi)
print("tile period (issue done -> issue done)      ", m(mma[lo + 1:hi + 1, 2] - mma[lo:hi, 2]))
print("MMA warp: wait for acc_empty                 ", m(mma[T, 1] - mma[T, 0]))
print("MMA warp: issue of one tile (16 k-steps)     ", m(mma[T, 2] - mma[T, 1]))
ev = T[T % 2 == 0]; od = T[T % 2 == 1]
for name, e, sel in (("warp 4 (group 0)", e4, ev), ("warp 11 (group 0)", e11, ev), ("warp 12 (group 1)", e12, od)):
    print(f"{name}: wait acc_full {m(e[sel, 1] - e[sel, 0])} | math {m(e[sel, 2] - e[sel, 1])} | "
          f"tile issued -> woken {m(e[sel, 1] - mma[sel, 2])} | math done -> MMA passes acc_empty of tile+2 "
          f"{m(mma[sel + 2, 1] - e[sel, 2])}")
print("first 24 tiles (cycles since first event): issue-start, issue-end | epi wake, epi done")
t0 = mma[0, 0]
for i in range(min(24, n)):
    e = e4 if i % 2 == 0 else e12
    print(f"  tile {i:3d}: {mma[i,1]-t0:8d} {mma[i,2]-t0:8d} | {e[i,1]-t0:8d} {e[i,2]-t0:8d}")
