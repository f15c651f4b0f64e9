"""Event trace of one CTA of the backward tensor kernel (library built with `make trace`, MSCS_LIB set)."""
import sys, os, ctypes as C
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
lib.mscs_debug_trace_bwd(buf.ctypes.data, 8192)
step()
lib.mscs_debug_trace_bwd(buf.ctypes.data, 8192)
tr = buf.astype(np.int64).reshape(4, 256, 8)
mma = tr[0]
arr = np.concatenate([tr[1], tr[2]], axis=1)       # [tile][16 warps] arrival clocks
m = lambda x: int(np.mean(x))
R = slice(40, 200)
print("tile period                            ", m(mma[41:200, 2] - mma[40:199, 2]))
print("MMA: dX(j-1) issued -> S(j+1) issued   ", m(mma[41:200, 0] - mma[40:199, 2]))
print("MMA: S(j+1) issued -> w_full(j) passed ", m(mma[R, 1] - mma[R, 0]))
print("MMA: w_full passed -> dX issued        ", m(mma[R, 2] - mma[R, 1]))
print("last arrival -> MMA w_full passed      ", m(mma[R, 1] - arr[R].max(axis=1)))
print("arrival spread (last - first)          ", m(arr[R].max(axis=1) - arr[R].min(axis=1)))
print("S(j) issued [tile j-1] -> first arrival", m(arr[41:200].min(axis=1) - mma[40:199, 0]))
print("S(j) issued [tile j-1] -> last arrival ", m(arr[41:200].max(axis=1) - mma[40:199, 0]))
print("mean arrival offset per warp (vs first):", [m(arr[R, w] - arr[R].min(axis=1)) for w in range(16)])
last = arr[R].argmax(axis=1)
print("how often each warp is last:", np.bincount(last, minlength=16).tolist())
seg = tr[3]
k0, k1 = seg[255, 7], seg[255, 6]
print("kernel (CTA 5):", k1 - k0, "cycles")
for i in range(12):
    if seg[i, 0]:
        print(f"run {i}: start +{seg[i,0]-k0}  tiles {seg[i,2]}  all MMAs issued +{seg[i,1]-k0}  -> {(seg[i,1]-seg[i,0])/max(1,seg[i,2]):.0f} cycles/tile")
flat = buf.astype(np.int64)
dur = flat[3 * 2048 + 1024: 3 * 2048 + 1024 + 148]
t_start = flat[3 * 2048 + 1280: 3 * 2048 + 1280 + 148]
t_end = flat[3 * 2048 + 1536: 3 * 2048 + 1536 + 148]
print("per-CTA cycles: min", dur.min(), "median", int(np.median(dur)), "max", dur.max(), "argmax", int(dur.argmax()))
print("start spread (ns):", t_start.max() - t_start.min(), " kernel span (ns):", t_end.max() - t_start.min(),
      " end spread (ns):", t_end.max() - t_end.min())

print("run boundaries seen by epilogue warp 4 (cycles since kernel start):")
for i in range(1, 6):
    b0, b1 = seg[64 + i - 1], seg[64 + i]
    if b1[3]:
        print(f" end of run {i-1}: last tile done +{b0[0]-k0} | df_full +{b0[1]-b0[0]} | flush +{b0[2]-b0[1]} | next() +{b1[3]-b0[2]} | X in TMEM +{b1[4]-b1[3]}"
              f" | row info + coefs +{b1[5]-b1[4]} | MMA run start at {seg[i,0]-k0} (= +{seg[i,0]-b1[4]} after X)")
