"""Single GPU: the forward sweeps of cfg5 restricted to the anchor-row range of rank r of `world` (what a pooled rank
runs), timed with CUDA events, next to the unrestricted sweeps -- separates the cost of row-range sharding itself from
multi-GPU effects.  usage: python tools/sweep_range_probe.py [world]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mscs_b200
from mscs_b200 import synth, _lib, _ops
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["cfg5"]
labels = synth.make_labels(cfg).to(dev)
fts = [torch.randn(cfg["n"], cfg["C"], cfg["H"] // s, cfg["W"] // s, device=dev) for s in cfg["strides"]]
mod = mscs_b200.DenseContrastiveLossV2_ms(dict(cfg["loss"]))
torch.manual_seed(0)
with torch.no_grad():
    mod(labels, fts)
torch.cuda.synchronize()
state = mod.last_state
job, sp = state.job, state.sp
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
buf = np.zeros(8, np.float32)
def run(tag):
    acc = np.zeros(8); K = 5
    for _ in range(K):
        state.stats.zero_()
        _lib.check(lib.mscs_sim_forward_sweeps(C.byref(job), st), "sweeps")
        k = lib.mscs_debug_fwd_timeline(buf.ctypes.data, 8)
        acc[:k] += buf[:k]
    print(tag, "row_ranges, work tables, sweep0, sweep1 (us):", [round(float(x) / K * 1e3, 1) for x in acc[:k]], flush=True)
run("all rows        ")
for r in (0, world // 2, world - 1):
    for i, (a, k, *_rest) in enumerate(sp.terms):
        t = job.terms[i]
        t.row_begin, t.row_end = _ops.shard_rows(t.N1, world, r)
        t.krow_begin, t.krow_end = _ops.shard_rows(t.N2, world, r)
    run(f"rank {r} of {world}     ")
