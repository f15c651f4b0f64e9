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


def test_philox_known_answers_and_stream():
    """Philox4x32-10 restatement (oracle/philox.py) against the known-answer vectors of the Random123 distribution
    (kat_vectors: zero, all-ones and pi-digit counters/keys), and the stream layout used by the opt-in sampler."""
    from oracle.philox import PhiloxStream, philox4x32_10, stream_words
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10([np.uint32(c) for c in ctr], [np.uint32(k) for k in key])
        assert tuple(int(g) for g in got) == want
    seed, call = 0x1234567887654321, 7
    w = stream_words(seed, call, 0, 64)
    blk3 = philox4x32_10([np.uint32(3), np.uint32(0), np.uint32(call), np.uint32(0)],
                         [np.uint32(seed & 0xffffffff), np.uint32(seed >> 32)])
    assert [int(x) for x in w[12:16]] == [int(x) for x in blk3]
    assert np.array_equal(stream_words(seed, call, 5, 20), w[5:25])
    g = PhiloxStream(seed, call)
    assert np.array_equal(np.concatenate([g.draw(3), g.draw(10), g.draw(51)]), w)
    assert not np.array_equal(stream_words(seed, call + 1, 0, 64), w)
    # the sampling oracle runs unchanged on this generator (same draw(k) interface as MT19937)
    from oracle import sampling
    lab = np.random.default_rng(0).integers(0, 4, (2, 16, 32)).astype(np.int64)
    a = sampling.sample_indices(lab, 16, 5, 2, 8, 1000, PhiloxStream(1, 0))
    b = sampling.sample_indices(lab, 16, 5, 2, 8, 1000, PhiloxStream(1, 0))
    c = sampling.sample_indices(lab, 16, 5, 2, 8, 1000, PhiloxStream(1, 1))
    assert np.array_equal(a["idx"], b["idx"]) and not np.array_equal(a["idx"], c["idx"])
