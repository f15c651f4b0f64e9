"""Per-CTA spans of the forward sweeps at one configuration (profiling build: `make prof`, MSCS_LIB=.../libmscs_prof.so):
how evenly the static work partition loads the persistent CTAs.  usage: python tools/cta_spans.py [cfg2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mscs_b200
from mscs_b200 import synth, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = torch.device("cuda:0")
cfg = synth.CONFIGS[name]
labels, feats = synth.make_inputs(name)
cls = mscs_b200.DenseContrastiveLossV2 if cfg["single_scale"] else mscs_b200.DenseContrastiveLossV2_ms
mod = cls(dict(cfg["loss"]))
labels = labels.to(dev); fg = [f.to(dev).requires_grad_(True) for f in feats]
torch.manual_seed(0)
def step():
    for f in fg: f.grad = None
    mod(labels, fg[0] if cfg["single_scale"] else fg).backward()
for _ in range(5): step()
torch.cuda.synchronize()
lib = _lib.load()
spans = np.zeros((160, 4), np.uint64)
for mode in (0, 1):
    if lib.mscs_debug_cta_spans_fwd(spans.ctypes.data, mode) == 0:
        print("not a profiling build"); break
    sp = spans[:160].astype(np.int64)
    sp = sp[sp[:, 1] > 0]
    t0, t1, cyc, sm = sp[:, 0], sp[:, 1], sp[:, 2], sp[:, 3]
    dur = (t1 - t0) / 1e3
    print(f"{name} sweep{mode}: {len(sp)} CTAs on {len(set(sm.tolist()))} SMs | kernel span {(t1.max() - t0.min()) / 1e3:.1f} us | start spread "
          f"{(t0.max() - t0.min()) / 1e3:.1f} us | duration min/median/max {dur.min():.1f}/{np.median(dur):.1f}/{dur.max():.1f} us | "
          f"sorted durations (every 10th): {[round(float(x), 1) for x in np.sort(dur)[::10]]}", flush=True)
