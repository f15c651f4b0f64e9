"""Barrier wait profile of the tensor kernels over a few cfg2 steps (run on the GPU box)."""
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
ns = np.zeros(32, np.uint64); cnt = np.zeros(32, np.uint64)
for which in ("fwd", "bwd"):
    getattr(lib, f"mscs_debug_wait_profile_{which}")(ns.ctypes.data, cnt.ctypes.data)
K = 5
for _ in range(K): step()
names = {101 % 32: "fwd producer a_empty", 102 % 32: "fwd producer b_empty", 111 % 32: "fwd mma a_full", 112 % 32: "fwd mma acc_empty",
         113 % 32: "fwd mma b_full", 121 % 32: "fwd epi acc_full(x8 warps x32 lanes)"}
namesb = {201 % 32: "bwd producer a_empty", 202 % 32: "bwd producer b_empty", 211 % 32: "bwd mma b_full", 212 % 32: "bwd mma a_full",
          213 % 32: "bwd mma df_empty", 214 % 32: "bwd mma w_full0", 215 % 32: "bwd mma w_full1", 221 % 32: "bwd epi s_full (x256 thr)", 222 % 32: "bwd epi df_full (x256 thr)"}
for which, nm in (("fwd", names), ("bwd", namesb)):
    getattr(lib, f"mscs_debug_wait_profile_{which}")(ns.ctypes.data, cnt.ctypes.data)
    print(which, "(per step, summed over 148 CTAs; divide by 148 for per-CTA; epilogue tags are per thread)")
    if ns[31]:
        print(f"  CTA 0: {ns[31]/K/1e3:.1f} us/step in this kernel family, effective SM clock {cnt[31]/ns[31]:.3f} GHz")
    for t in range(31):
        if cnt[t]:
            print(f"  tag%32={t:2d} {nm.get(t,'?'):40s} waits/step {cnt[t]/K:10.0f}  us/step/CTA {ns[t]/K/148/1e3:10.1f}")
