"""Event trace of one CTA of sweep 1 (positive sweep) of the forward kernel (`make trace1`, MSCS_LIB=libmscs_trace1.so)."""
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
k_launch, k_go, k_end = tr[3, 255, 5], tr[3, 255, 7], tr[3, 255, 6]
print(f"tiles of CTA 5: {n}; prologue (TMEM alloc, barriers) -> dependency wait passed +{k_go - k_launch}; kernel body {k_end - k_go} cycles")
print("tile: [k_full wait start, passed] acc_empty passed, issue end | epilogue wake, done   (cycles since the dependency wait)")
for i in range(n):
    e = e4 if i % 2 == 0 else e12
    kf = f"{mma[i,3]-k_go:7d} {mma[i,4]-k_go:7d}" if mma[i, 3] else "      -       -"
    print(f"  {i:3d}: [{kf}] {mma[i,1]-k_go:7d} {mma[i,2]-k_go:7d} | {e[i,1]-k_go:7d} {e[i,2]-k_go:7d}")
