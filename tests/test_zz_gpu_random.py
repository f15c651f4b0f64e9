"""GPU parity on RANDOM configurations: the drop-in modules against the fp64 oracle on the cases of
`helpers.random_case` -- the same generator (same seed) the CPU test `test_oracle_vs_live_reference_cpu.py` walks
against the executable reference, so every case here is one on which the oracle is pinned.

`MSCS_GPU_RANDOM=<number of cases>` (default 10; 0 = skipped).  The first 10 cases ran green on a B200
(profiles/r01_gpu_random_cases.txt: 8 compared, 2 refused like the reference); larger counts walk cases that have
only been checked on the CPU side so far.

Tolerances (north_star): sampled indices / pair lists / generator state bit-exact; loss <= 1e-3 relative (+1e-5
absolute for one-class scales whose loss is ~0); gradients cosine >= 0.999.  The individual TERM losses (ms_losses /
cs_losses, logged only) get 3e-3 when the term has fewer than 256 anchor rows: the rounding of the bf16 operands
(2^-9 per component) does not average out over so few rows at C <= 48 and tau = 0.07 (case 60 of the 80-case walk on a
B200: one term of ~60 rows off by 1.01e-3 with the total inside 1e-3)."""
import os

import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
N_CASES = int(os.environ.get("MSCS_GPU_RANDOM", "10"))


@pytest.mark.skipif(N_CASES == 0, reason="MSCS_GPU_RANDOM=0")
def test_random_configs_vs_oracle():
    import mscs_b200
    from oracle import loss_fp64
    from oracle.mt19937 import MT19937
    assert torch.cuda.is_available() and mscs_b200.load().mscs_device_ok() == 1
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(20221017)          # the seed of the live-reference CPU test
    done = refused = 0
    for case in range(N_CASES):
        rc = helpers.random_case(rs)
        if rc is None:
            continue
        cfg, single, labels, feats, S = rc["cfg"], rc["single"], rc["labels"], rc["feats"], rc["S"]
        ocfg = helpers.oracle_cfg_for(dict(loss_cfg=cfg, single_scale=single))
        torch.manual_seed(rc["seed"])
        state0 = torch.get_rng_state()
        gen = MT19937.from_torch_state(state0.numpy().tobytes())
        try:
            want = loss_fp64.ms_cs_loss(labels.numpy(), [f.numpy() for f in feats], ocfg, gen, need_grad=True)
        except Exception:
            want = None
        mod = helpers.make_module(dict(loss_cfg=cfg, single_scale=single))
        fg = [f.to(dev).requires_grad_(True) for f in feats]
        if want is None or not np.isfinite(want["total"]):
            if want is None:       # Q8: no kept pair (RuntimeError) or a one-pixel class (IndexError), like the reference
                with pytest.raises((RuntimeError, IndexError)):
                    mod(labels.to(dev), fg[0] if single else fg)
                refused += 1
                print(f"refused case {case} like the reference", flush=True)
            continue
        loss = mod(labels.to(dev), fg[0] if single else fg)
        loss.backward()
        tag = f"case {case}: {cfg} strides {rc['strides']} n {labels.shape[0]} C {feats[0].shape[1]}"
        assert abs(float(loss) - want["total"]) <= 1e-3 * abs(want["total"]) + 1e-5, (tag, float(loss), want["total"])
        if not single:
            got_terms = [float(x) for x in list(mod.ms_losses) + list(mod.cs_losses)]
            assert len(got_terms) == len(want["ms"]) + len(want["cs"]), tag
            rows_of = [sm["T"] * sm["V"] for sm in want["samples"]]
            n_rows = rows_of + [rows_of[0]] * len(want["cs"])          # cs terms: rows = scale-0 anchors
            for a, b, nr in zip(got_terms, want["ms"] + want["cs"], n_rows):
                assert abs(a - b) <= (1e-3 if nr >= 256 else 3e-3) * abs(b) + 1e-5, (tag, a, b, nr)
        # the generator ends where the reference's randperm calls leave it
        after = MT19937.from_torch_state(torch.get_rng_state().numpy().tobytes())
        assert gen.pos == after.pos and np.array_equal(gen.mt, after.mt), tag
        for s in range(S):
            g64 = want["grads"][s]
            got = fg[s].grad.cpu().numpy().astype(np.float64) if fg[s].grad is not None else np.zeros_like(g64)
            # bit-exact sampled set: the gradient is non-zero exactly at the oracle's sampled pixels (rows of a
            # one-class scale carry a zero gradient on both sides and are skipped)
            if max(np.abs(g64).max(), np.abs(got).max()) < 1e-7:
                continue
            assert helpers.cosine(got, g64) >= 0.999, (tag, s, helpers.cosine(got, g64))
            touched = np.abs(got).sum(1) != 0
            smp = want["samples"][s]
            ref_mask = np.zeros(touched.shape, dtype=bool).reshape(touched.shape[0], -1)
            for k in range(smp["idx"].shape[0]):
                ref_mask[smp["pairs"][k, 0], smp["idx"][k]] = True
            assert not (touched.reshape(ref_mask.shape) & ~ref_mask).any(), (tag, s, "gradient outside the sampled set")
        done += 1
        print(f"ok {tag}: loss {float(loss):.6f} oracle {want['total']:.6f}", flush=True)
    print(f"{done} of {N_CASES} random cases compared on the GPU, {refused} refused like the reference")
    assert done >= N_CASES // 2
