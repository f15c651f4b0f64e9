"""CPU: the oracle against the fixtures generated from the executable reference
(tests/golden/make_golden.py) and against torch's own randperm / nearest interpolation."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import load_npz, oracle_cfg_for, small_case_inputs, cosine
from oracle.mt19937 import MT19937, randperm_full, randperm_prefix
from oracle.sampling import nearest_source_index, sample_indices
from oracle import loss_fp64, torch_port

SMALL = ["tiny_ss", "tiny_ms", "tiny_ms_detach", "odd_ss"]


def test_mt19937_matches_numpy_and_torch():
    g = MT19937(seed=123)
    ref = np.random.RandomState(123).randint(0, 2 ** 32, size=3000, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(g.draw(3000), ref)
    torch.manual_seed(77)
    g = MT19937.from_torch_state(torch.get_rng_state().numpy().tobytes())
    assert np.array_equal(randperm_full(1000, g), torch.randperm(1000).numpy())
    # mid-block state round trip
    g2 = MT19937.from_torch_state(torch.get_rng_state().numpy().tobytes())
    assert g2.pos == g.pos and np.array_equal(g2.mt, g.mt)


def test_randperm_known_answers():
    z = load_npz("aten_known_answers")
    for key in z.files:
        if not key.startswith("randperm_s"):
            continue
        seed, n = int(key.split("_s")[1].split("_")[0]), int(key.split("_n")[1])
        assert np.array_equal(randperm_full(n, MT19937(seed=seed)), z[key]), key
        assert np.array_equal(randperm_prefix(n, min(n, 37), MT19937(seed=seed)), z[key][:37]), key
    g = MT19937(seed=5)
    chain = np.concatenate([randperm_full(n, g) for n in (3, 700, 2, 625, 1249)])
    assert np.array_equal(chain, z["randperm_chain"])


def test_nearest_known_answers():
    z = load_npz("aten_known_answers")
    for key in z.files:
        if not key.startswith("nearest_"):
            continue
        hw, fw = key.split("_")[1], int(key.split("fw")[1])
        H, W = (int(v) for v in hw.split("x"))
        s = W // fw
        lab = (np.arange(H * W).reshape(H, W) % 1000)
        ys, xs = nearest_source_index(H // s, H), nearest_source_index(W // s, W)
        assert np.array_equal(lab[ys][:, xs], z[key]), key


@pytest.mark.parametrize("name", SMALL)
def test_small_cases_full(name, golden):
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    cfg = oracle_cfg_for(meta)
    gen = MT19937.from_torch_state(z["rng_state0"].tobytes())
    res = loss_fp64.ms_cs_loss(labels.numpy(), [f.numpy() for f in feats], cfg, gen)
    for s in range(len(feats)):
        assert np.array_equal(res["samples"][s]["idx"], z[f"idx{s}"])
        assert np.array_equal(res["samples"][s]["pairs"], z[f"pairs{s}"])
        assert cosine(res["grads"][s], z[f"grad{s}"]) > 0.999999
        assert np.abs(res["grads"][s] - z[f"grad{s}"]).max() < 1e-6
    after = MT19937.from_torch_state(z["rng_state1"].tobytes())
    assert gen.pos == after.pos and np.array_equal(gen.mt, after.mt)
    assert abs(res["total"] - meta["total"]) < 2e-6 * abs(meta["total"])
    for a, b in zip(res["ms"] + res["cs"], meta["ms"] + meta["cs"]):
        assert abs(a - b) < 2e-6 * abs(b)
    assert len(res["cs"]) == len(meta["cs"])


@pytest.mark.parametrize("name", SMALL)
def test_torch_port_matches_reference(name, golden):
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    cfg = oracle_cfg_for(meta)
    torch.set_rng_state(torch.from_numpy(z["rng_state0"]))
    fg = [f.clone().requires_grad_(True) for f in feats]
    total, ms, cs, idx = torch_port.ms_cs_loss(labels, fg, cfg)
    total.backward()
    assert abs(float(total) - meta["total"]) < 1e-6 * abs(meta["total"])
    for s in range(len(feats)):
        assert np.array_equal(idx[s].numpy(), z[f"idx{s}"])
        assert np.abs(fg[s].grad.numpy() - z[f"grad{s}"]).max() < 1e-6
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg4_large"])
def test_big_sampling_hashes(name, golden):
    """(T,V) and index hashes of the benchmark configurations (sampling only: cheap on CPU)."""
    from mscs_b200 import synth
    from oracle.config import oracle_cfg
    from helpers import CLASSES
    meta = golden[name + "_sampling"]
    cfg = synth.CONFIGS[name]
    labels, _ = synth.make_inputs(name, with_features=False)
    lc = cfg["loss"]
    ocfg = oracle_cfg(lc, CLASSES[(lc["dataset"], lc["experiment"])])
    torch.manual_seed(0)
    gen = MT19937.from_torch_state(torch.get_rng_state().numpy().tobytes())
    for s, stride in enumerate(cfg["strides"]):
        o = sample_indices(labels.numpy(), cfg["W"] // stride, ocfg["num_all_classes"], ocfg["min_views"],
                           ocfg["max_views"], ocfg["max_total"], gen)
        assert [o["T"], o["V"]] == meta["TV"][s]
        assert hashlib.sha256(np.ascontiguousarray(o["idx"].astype(np.int64)).tobytes()).hexdigest() == meta["idx_sha"][s]
