"""Host-side profile of the step loop (cProfile) -- run on the GPU box."""
import cProfile
import pstats
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mscs_b200
from mscs_b200 import synth

dev = torch.device("cuda:0")
cfg = synth.CONFIGS["cfg2"]
labels, feats = synth.make_inputs("cfg2")
labels = labels.to(dev)
feats = [f.to(dev).requires_grad_(True) for f in feats]
mod = mscs_b200.DenseContrastiveLossV2_ms(dict(cfg["loss"]))
torch.manual_seed(0)


def step():
    for f in feats:
        f.grad = None
    loss = mod(labels, feats)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
