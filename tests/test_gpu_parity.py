"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle and the golden fixtures generated from the executable reference.

Tolerances (BASELINE.json north_star): sampled indices / pair lists / T / V bit-exact; loss within
1e-3 relative; embedding gradients cosine >= 0.999 (max-abs error printed)."""
import ctypes as C
import hashlib

import numpy as np
import pytest
import torch

from helpers import CLASSES, cosine, load_npz, make_module, oracle_cfg_for, small_case_inputs

pytestmark = pytest.mark.gpu

SMALL = ["tiny_ss", "tiny_ms", "tiny_ms_detach", "odd_ss"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import mscs_b200
    assert mscs_b200.load().mscs_device_ok() == 1, "libmscs.so needs a compute-capability 10.x device"
    return torch.device("cuda:0")


def _sample(meta, labels, feats, z, dev):
    from mscs_b200 import _ops
    mod = make_module(meta)
    torch.set_rng_state(torch.from_numpy(z["rng_state0"]))
    smp = _ops.sample_anchors(labels.to(dev), [tuple(f.shape[-2:]) for f in feats], mod._spec)
    return mod, smp


def _check_sorted_layout(smp, idx_ref, pairs, plane, A):
    """The class-sorted kernel layout must hold exactly the reference's sampled set."""
    pix, cls, seg = smp.pix.cpu().numpy(), smp.cls.cpu().numpy(), smp.seg.cpu().numpy()
    ref_pix = (pairs[:, 0:1].astype(np.int64) * plane + idx_ref).ravel()
    ref_cls = np.repeat(pairs[:, 1], idx_ref.shape[1])
    assert np.array_equal(np.sort(pix), np.sort(ref_pix))
    assert len(np.unique(pix)) == len(pix)
    assert np.all(np.diff(cls) >= 0)
    order = np.argsort(ref_pix)
    assert np.array_equal(cls[np.argsort(pix)], ref_cls[order])
    counts = np.bincount(cls, minlength=A)
    assert np.array_equal(seg, np.concatenate([[0], np.cumsum(counts)]))


@pytest.mark.parametrize("name", SMALL)
def test_sampling_bit_exact_small(name, golden, dev):
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    mod, smp = _sample(meta, labels, feats, z, dev)
    for s, sm in enumerate(smp):
        assert [sm.T, sm.V] == meta["TV"][s]
        assert np.array_equal(sm.idx_ref.cpu().numpy(), z[f"idx{s}"]), f"scale {s}: sampled indices differ"
        assert np.array_equal(sm.pair_ref.cpu().numpy(), z[f"pairs{s}"])
        _check_sorted_layout(sm, z[f"idx{s}"], z[f"pairs{s}"], sm.dl_h * sm.dl_w, mod._spec.num_classes)
    # the torch CPU generator must end where the reference leaves it
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_sampling_bit_exact_bench_configs(name, golden, dev):
    from mscs_b200 import synth
    meta = golden[name]
    z = load_npz(name)
    labels, _ = synth.make_inputs(name, with_features=(name == "cfg1"))
    cfg = synth.CONFIGS[name]
    hw = [(cfg["H"] // s, cfg["W"] // s) for s in cfg["strides"]]
    from mscs_b200 import _ops
    mod = make_module(meta)
    torch.set_rng_state(torch.from_numpy(z["rng_state0"]))
    smp = _ops.sample_anchors(labels.to(dev), hw, mod._spec)
    for s, sm in enumerate(smp):
        assert [sm.T, sm.V] == meta["TV"][s]
        assert np.array_equal(sm.idx_ref.cpu().numpy(), z[f"idx{s}"])
        assert np.array_equal(sm.pair_ref.cpu().numpy(), z[f"pairs{s}"])
        _check_sorted_layout(sm, z[f"idx{s}"], z[f"pairs{s}"], sm.dl_h * sm.dl_w, mod._spec.num_classes)
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg4_large", "cfg5"])
def test_sampling_hashes_big(name, golden, dev):
    import mscs_b200
    from mscs_b200 import _ops, synth
    meta = golden[name + "_sampling"]
    z = load_npz(name + "_sampling")
    cfg = synth.CONFIGS[name]
    labels, _ = synth.make_inputs(name, with_features=False)
    cls = mscs_b200.DenseContrastiveLossV2 if cfg["single_scale"] else mscs_b200.DenseContrastiveLossV2_ms
    mod = cls(dict(cfg["loss"]))
    torch.manual_seed(0)
    smp = _ops.sample_anchors(labels.to(dev), [(cfg["H"] // s, cfg["W"] // s) for s in cfg["strides"]], mod._spec)
    for s, sm in enumerate(smp):
        assert [sm.T, sm.V] == meta["TV"][s]
        idx = sm.idx_ref.cpu().numpy().astype(np.int64)
        assert hashlib.sha256(np.ascontiguousarray(idx).tobytes()).hexdigest() == meta["idx_sha"][s]
        assert np.array_equal(sm.pair_ref.cpu().numpy(), z[f"pairs{s}"])
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


def test_sampling_many_views_single_image(dev):
    """Validation-like shape: one image, few pairs, thousands of views per class (exercises the
    Fisher-Yates prefix resolution with many collisions) against the oracle."""
    from mscs_b200 import _ops, synth
    from oracle.mt19937 import MT19937
    from oracle.sampling import sample_indices
    labels = synth.synth_labels(1, 256, 512, 19, 5, 16, 0.05, 9)
    spec = _ops.LossSpec(num_classes=20, temperature=0.1, cs_temperature=0.1, min_views=5, max_views=2500,
                         max_total=10000)
    for seed, fw in [(3, 128), (4, 512)]:
        torch.manual_seed(seed)
        gen = MT19937.from_torch_state(torch.get_rng_state().numpy().tobytes())
        o = sample_indices(labels.numpy(), fw, 20, 5, 2500, 10000, gen)
        sm = _ops.sample_anchors(labels.to(dev), [(256 * fw // 512, fw)], spec)[0]
        assert (sm.T, sm.V) == (o["T"], o["V"]) and sm.V >= 1000
        assert np.array_equal(sm.idx_ref.cpu().numpy(), o["idx"])


def test_sampling_errors(dev):
    from mscs_b200 import _ops
    spec = _ops.LossSpec(num_classes=20, temperature=0.1, cs_temperature=0.1, min_views=5)
    labels = torch.full((1, 32, 32), 19, dtype=torch.long, device=dev)     # ignore class only
    with pytest.raises(RuntimeError):
        _ops.sample_anchors(labels, [(8, 8)], spec)


def _sets_for(meta, labels, feats, z, dev):
    from mscs_b200 import _ops
    mod, smp = _sample(meta, labels, feats, z, dev)
    sets = [_ops.gather_normalize(f.to(dev).contiguous(), s) for f, s in zip(feats, smp)]
    return mod, smp, sets


@pytest.mark.parametrize("name", SMALL)
def test_gather_normalize(name, golden, dev):
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    mod, smp, sets = _sets_for(meta, labels, feats, z, dev)
    for f, sm, st in zip(feats, smp, sets):
        n, Cc, h, w = f.shape
        pix = sm.pix.cpu().long()
        x = f.reshape(n, Cc, -1)[pix // (h * w), :, pix % (h * w)]
        ref = torch.nn.functional.normalize(x, p=2, dim=1)
        assert torch.allclose(st.f32.cpu(), ref, atol=2e-7, rtol=1e-6)
        assert torch.allclose(st.inv_norm.cpu(), 1.0 / x.norm(dim=1).clamp_min(1e-12), rtol=1e-6)
        bf = st.bf16.cpu()
        assert torch.equal(bf[:st.N, :Cc], st.f32.cpu().bfloat16())       # same rounding of the same fp32 value
        assert bf[st.N:].abs().max() == 0 and (Cc == st.C_pad or bf[:, Cc:].abs().max() == 0)


def _oracle_terms(spec, smp, rows, single_scale):
    """fp64 oracle on given unit rows (one matrix per scale, kernel row order)."""
    from oracle import loss_fp64
    ys = [s.cls.cpu().numpy() for s in smp]
    S = len(rows)
    terms = [(s, s, True, spec.weights[s], spec.temperature, False) for s in range(S)]
    if spec.cross_scale and not single_scale:
        terms.append((0, S - 1, False, spec.w_high_low, spec.cs_temperature, not spec.detach_deepest))
        if S > 2:
            terms.append((0, S - 2, False, spec.w_high_mid, spec.cs_temperature, not spec.detach_deepest))
    out, dFs, total = [], [np.zeros_like(r) for r in rows], 0.0
    for (a, k, sm_, w, tau, need_dk) in terms:
        l, da, dk, st = loss_fp64.term(rows[a], ys[a], rows[k], ys[k], tau, sm_, True, 512)
        out.append((l, st))
        total += w * l
        dFs[a] += w * da
        if sm_ or need_dk:
            dFs[k] += w * dk
    return out, dFs, total


def _pad_f32(st, src):
    out = torch.zeros((st.N, st.C_pad), dtype=torch.float32, device=src.device)
    out[:, :src.shape[1]] = src[:st.N, :src.shape[1]]
    return out


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("name", SMALL)
def test_similarity_kernels_vs_oracle(name, impl, golden, dev):
    """Row statistics, per-term losses and d loss/d unit rows of the similarity kernels against the
    fp64 oracle evaluated on the SAME operand rows (fp32 rows for the SIMT validation kernels, the
    bf16-rounded rows for the tcgen05 kernels), so only accumulation order / exp2 / bf16 W remain."""
    from mscs_b200 import _lib, _ops
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    mod, smp, sets = _sets_for(meta, labels, feats, z, dev)
    spec, single = mod._spec, meta["single_scale"]
    state = _ops.build_job(spec, smp, sets, single)
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if impl == "simt":
        f32 = [_pad_f32(s, s.f32) for s in sets]
        rows = [s.f32.cpu().double().numpy() for s in sets]
    else:
        f32 = None
        rows = [s.bf16[:s.N, :s.C].float().cpu().double().numpy() for s in sets]
    g = torch.full((1,), 0.7, dtype=torch.float32, device=dev)
    dFs = [torch.zeros((s.N, s.C_pad), dtype=torch.float32, device=dev) for s in sets]
    ptrs = [0] * _lib.MAX_SCALES
    lds = (C.c_int32 * _lib.MAX_SCALES)()
    for i, d in enumerate(dFs):
        ptrs[i], lds[i] = d.data_ptr(), d.shape[1]
    if impl == "simt":
        fp = [0] * _lib.MAX_SCALES
        for i, f in enumerate(f32):
            fp[i] = f.data_ptr()
        _lib.check(lib.mscs_debug_sim_forward_simt(C.byref(state.job), _lib.ptr_array(fp), st), "simt fwd")
        _lib.check(lib.mscs_debug_sim_backward_simt(C.byref(state.job), _lib.ptr_array(fp), g.data_ptr(),
                                                    _lib.ptr_array(ptrs), lds, st), "simt bwd")
    else:
        _lib.check(lib.mscs_sim_forward(C.byref(state.job), st), "tc fwd")
        _lib.check(lib.mscs_sim_backward(C.byref(state.job), g.data_ptr(), _lib.ptr_array(ptrs), lds, st), "tc bwd")
    torch.cuda.synchronize()
    terms, o_dFs, o_total = _oracle_terms(spec, smp, rows, single)
    stats = state.keep[0].cpu().numpy()
    off = 0
    for i, (l, ost) in enumerate(terms):
        n1 = len(ost["neg"])
        neg, pos, ssum = stats[off:off + n1], stats[off + n1:off + 2 * n1], stats[off + 2 * n1:off + 3 * n1]
        off += 3 * n1
        np.testing.assert_allclose(neg, ost["neg"], rtol=2e-5, err_msg=f"{impl} term {i} neg sums")
        np.testing.assert_allclose(pos, ost["possum"], rtol=1e-4, atol=1e-4, err_msg=f"{impl} term {i} pos sums")
        np.testing.assert_allclose(ssum, ost["S"], rtol=1e-4, atol=1e-9, err_msg=f"{impl} term {i} S sums")
        assert abs(float(state.term_loss[i]) - l) < 2e-5 * abs(l)
    assert abs(float(state.total) - o_total) < 2e-5 * abs(o_total)
    for s in range(len(sets)):
        got = dFs[s][:, :sets[s].C].cpu().double().numpy()
        want = 0.7 * o_dFs[s]
        cs = cosine(got, want)
        err = np.abs(got - want).max()
        print(f"{name}/{impl} set {s}: dF cosine {cs:.8f} max-abs {err:.3e} (scale {np.abs(want).max():.3e})")
        assert cs > (0.999999 if impl == "simt" else 0.9999)
        assert err < (1e-5 if impl == "simt" else 2e-2) * np.abs(want).max()


@pytest.mark.parametrize("name", SMALL)
def test_end_to_end_small(name, golden, dev):
    """Module forward/backward (the drop-in call) against the reference's recorded outputs."""
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    mod = make_module(meta)
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    torch.set_rng_state(torch.from_numpy(z["rng_state0"]))
    loss = mod(labels.to(dev), fg[0] if meta["single_scale"] else fg)
    assert loss.dim() == 0 and loss.requires_grad
    loss *= 0.1                       # LossWrapper multiplies in place (LossWrapper.py:90)
    loss.backward()
    assert abs(float(loss) / 0.1 - meta["total"]) < 1e-3 * abs(meta["total"])
    if not meta["single_scale"]:
        assert len(mod.ms_losses) == len(meta["ms"]) and len(mod.cs_losses) == len(meta["cs"])
        for a, b in zip(list(mod.ms_losses) + list(mod.cs_losses), meta["ms"] + meta["cs"]):
            assert abs(float(a) - b) < 1e-3 * abs(b)
    for s, f in enumerate(fg):
        got, want = f.grad.cpu().numpy() / 0.1, z[f"grad{s}"]
        cs, err = cosine(got, want), np.abs(got - want).max()
        print(f"{name} scale {s}: grad cosine {cs:.7f} max-abs {err:.3e} (ref max {np.abs(want).max():.3e})")
        assert cs >= 0.999
        assert np.array_equal(got != 0, want != 0) or np.abs(got[(got != 0) != (want != 0)]).max() < 1e-12
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg4_large", "cfg5"])
def test_end_to_end_bench_configs(name, golden, dev):
    """Every BASELINE configuration against recorded loss / gradient rows: cfg-1..4 from the REFERENCE run on the build
    machine, cfg-4-large and cfg-5 (64 images on one GPU) from the chunked fp64 oracle, which the other cases pin
    (tests/golden/make_golden.py; golden.json says which under `source`)."""
    from mscs_b200 import synth
    meta = golden[name]
    z = load_npz(name)
    labels, feats = synth.make_inputs(name)
    mod = make_module(meta)
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    torch.set_rng_state(torch.from_numpy(z["rng_state0"]))
    loss = mod(labels.to(dev), fg[0] if meta["single_scale"] else fg)
    loss.backward()
    rel = abs(float(loss) - meta["total"]) / abs(meta["total"])
    print(f"{name}: loss {float(loss):.6f} reference {meta['total']:.6f} rel {rel:.2e}")
    assert rel < 1e-3
    if not meta["single_scale"]:
        for a, b in zip(list(mod.ms_losses) + list(mod.cs_losses), meta["ms"] + meta["cs"]):
            assert abs(float(a) - b) < 1e-3 * abs(b)
    for s, f in enumerate(fg):
        n, Cc, h, w = f.shape
        g = f.grad.reshape(n, Cc, h * w)
        idx, pairs = z[f"idx{s}"], z[f"pairs{s}"]
        T, V = idx.shape
        ids = z[f"grad_row_ids{s}"]
        k, v = ids // V, ids % V
        bsel = torch.from_numpy(pairs[k, 0].astype(np.int64)).to(dev)
        psel = torch.from_numpy(idx[k, v].astype(np.int64)).to(dev)
        rows = g[bsel, :, psel].cpu().numpy()
        cs, err = cosine(rows, z[f"grad_rows{s}"]), np.abs(rows - z[f"grad_rows{s}"]).max()
        # per-row L2 norms of every sampled pixel + global norm
        ball = torch.from_numpy(np.repeat(pairs[:, 0], V).astype(np.int64)).to(dev)
        pall = torch.from_numpy(idx.ravel().astype(np.int64)).to(dev)
        l2 = g[ball, :, pall].norm(dim=1).cpu().numpy()
        l2_cos = cosine(l2, z[f"grad_row_l2_{s}"])
        tot = float(f.grad.double().norm())
        print(f"{name} scale {s}: grad rows cosine {cs:.7f} max-abs {err:.3e}; row-norm cosine {l2_cos:.7f}; "
              f"|grad| {tot:.6e} vs {meta['grad_l2'][s]:.6e}")
        assert cs >= 0.999 and l2_cos >= 0.999
        assert abs(tot - meta["grad_l2"][s]) < 1e-2 * meta["grad_l2"][s]
        assert int((f.grad != 0).any(dim=1).sum()) == T * V      # exactly the sampled pixels are touched


def test_properties_full_size(dev):
    """Size-independent properties at the headline size (cfg-2): the gradient of a function of the
    normalised embedding is orthogonal to the embedding; scaling a feature map by a power of two
    leaves the loss unchanged; same seed -> same sampled set."""
    import mscs_b200
    from mscs_b200 import synth
    labels, feats = synth.make_inputs("cfg2")
    mod = mscs_b200.DenseContrastiveLossV2_ms(dict(synth.CONFIGS["cfg2"]["loss"]))
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    torch.manual_seed(0)
    loss = mod(labels.to(dev), fg)
    loss.backward()
    pix0 = [s.pix.clone() for s in mod.last_samples]
    for f in fg:
        dot = (f.grad * f.detach()).sum(1)
        scale = (f.grad.norm(dim=1) * f.detach().norm(dim=1)).clamp_min(1e-30)
        assert float((dot.abs() / scale).max()) < 2e-3
    torch.manual_seed(0)
    with torch.no_grad():
        loss2 = mod(labels.to(dev), [f.detach() * 4.0 for f in fg])
    assert all(torch.equal(a, s.pix) for a, s in zip(pix0, mod.last_samples))
    assert abs(float(loss2) - float(loss)) < 1e-5 * abs(float(loss))


# ---- pooled cross-batch mode: `world` ranks emulated on one GPU (ThreadComm) ------------------
def _pooled_emulated(cfg, labels, feats, world, seed, dev, keep_on_device=False):
    import threading
    import mscs_b200
    from mscs_b200 import _ops
    n = labels.shape[0]
    nl = n // world
    shared = _ops.ThreadComm._Shared(world)
    results, errors = [None] * world, []

    def worker(r):
        try:
            torch.cuda.set_device(dev)
            mod = mscs_b200.DenseContrastiveLossV2_ms(dict(cfg), comm=_ops.ThreadComm(shared, r))
            lab = labels[r * nl:(r + 1) * nl].to(dev)
            fts = [f[r * nl:(r + 1) * nl].to(dev).requires_grad_(True) for f in feats]
            loss = mod(lab, fts)
            grads = _ops.run_backward(mod.last_state, torch.ones((), device=dev), [True] * len(fts),
                                      [tuple(f.shape) for f in fts], [f.dtype for f in fts])
            torch.cuda.synchronize()
            results[r] = (float(loss.detach()), [g if keep_on_device else g.cpu() for g in grads],
                          [float(x) for x in mod.ms_losses] + [float(x) for x in mod.cs_losses])
        except Exception as e:      # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    torch.manual_seed(seed)
    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.parametrize("world", [2, 4])
def test_pooled_mode_matches_single_process(world, dev):
    """The pooled loss over `world` ranks (rows sharded, keys exchanged) equals the single-process loss
    of the concatenated batch: same sampled pixels, same loss, same gradient for every rank's images."""
    import mscs_b200
    from mscs_b200 import synth
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, scales=3, weights=[1.0, 0.7, 0.4],
               cross_scale_contrast=True, w_high_low=0.5, w_high_mid=0.25, min_views_per_class=5,
               max_views_per_class=50, max_features_total=1500)
    n = 4
    labels = synth.synth_labels(n, 128, 256, 19, 7, 16, 0.05, 21)
    feats = synth.synth_features(n, 64, 128, 256, [4, 8, 16], 22)
    # single process, whole batch
    full = mscs_b200.DenseContrastiveLossV2_ms(dict(cfg))
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    torch.manual_seed(5)
    loss = full(labels.to(dev), fg)
    loss.backward()
    state_after = torch.get_rng_state()
    want_terms = [float(x) for x in full.ms_losses] + [float(x) for x in full.cs_losses]
    assert full.last_samples[0].N > 256          # several row tiles, so the row ranges matter
    res = _pooled_emulated(cfg, labels, feats, world, 5, dev)
    assert torch.equal(torch.get_rng_state(), state_after)
    nl = n // world
    for r, (l, grads, terms) in enumerate(res):
        assert abs(l - float(loss)) < 2e-5 * abs(float(loss)), (r, l, float(loss))
        for a, b in zip(terms, want_terms):
            assert abs(a - b) < 2e-5 * abs(b)
        for s, g in enumerate(grads):
            want = fg[s].grad[r * nl:(r + 1) * nl].cpu()
            assert torch.equal(g != 0, want != 0), f"rank {r} scale {s}: sampled pixels differ"
            cs = cosine(g.numpy(), want.numpy())
            err = float((g - want).abs().max())
            print(f"world {world} rank {r} scale {s}: grad cosine {cs:.8f} max-abs {err:.3e}")
            assert cs > 0.99999


@pytest.mark.parametrize("world", [2, 8])
def test_pooled_cfg5_vs_fp64_oracle_fixture(world, golden, dev):
    """BASELINE config 5 (64 images pooled over the ranks) through the POOLED code path -- `world` ranks emulated on
    this GPU (ThreadComm: same kernels, row ranges, peer-store gather, statistics push and pull scatter as under
    torchrun; only the transport is local) -- against the chunked fp64 oracle's loss and gradient rows
    (tests/golden/cfg5.npz).  tools/pooled_check.py repeats the comparison over real NVLink ranks."""
    from mscs_b200 import synth
    meta, z = golden["cfg5"], load_npz("cfg5")
    labels, feats = synth.make_inputs("cfg5")
    res = _pooled_emulated(dict(meta["loss_cfg"]), labels, feats, world, 0, dev, keep_on_device=True)
    nl = labels.shape[0] // world
    for r, (l, grads, terms) in enumerate(res):
        rel = abs(l - meta["total"]) / abs(meta["total"])
        assert rel < 1e-3, (r, l, meta["total"])
        for a, b in zip(terms, meta["ms"] + meta["cs"]):
            assert abs(a - b) < 1e-3 * abs(b), (r, a, b)
    print(f"pooled cfg5 x{world}: loss {res[0][0]:.6f} oracle {meta['total']:.6f}")
    for s in range(len(feats)):
        idx, pairs, ids = z[f"idx{s}"], z[f"pairs{s}"], z[f"grad_row_ids{s}"]
        T, V = idx.shape
        k, v = ids // V, ids % V
        rows = np.empty((len(ids), feats[s].shape[1]), dtype=np.float32)
        for j in range(len(ids)):
            b, p = int(pairs[k[j], 0]), int(idx[k[j], v[j]])
            g = res[b // nl][1][s]
            rows[j] = g[b % nl].reshape(g.shape[1], -1)[:, p].cpu().numpy()
        cs, err = cosine(rows, z[f"grad_rows{s}"]), np.abs(rows - z[f"grad_rows{s}"]).max()
        tot = float(np.sqrt(sum(float(res[r][1][s].double().pow(2).sum()) for r in range(world))))
        touched = sum(int((res[r][1][s] != 0).any(dim=1).sum()) for r in range(world))
        print(f"pooled cfg5 x{world} scale {s}: grad rows cosine {cs:.7f} max-abs {err:.3e}; |grad| {tot:.6e} vs "
              f"{meta['grad_l2'][s]:.6e}; {touched} pixels touched")
        assert cs >= 0.999
        assert abs(tot - meta["grad_l2"][s]) < 1e-2 * meta["grad_l2"][s]
        assert touched == T * V


# ---- API behaviours of the reference classes ----------------------------------------------------
def test_validation_shape_forward_only(dev):
    """validate() calls the loss under no_grad with n = 1 and full frames."""
    import mscs_b200
    from mscs_b200 import synth
    from oracle import loss_fp64
    from oracle.config import oracle_cfg
    from oracle.mt19937 import MT19937
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, scales=2, weights=[1.0, 0.5],
               cross_scale_contrast=True, min_views_per_class=5, max_views_per_class=300, max_features_total=2000)
    labels = synth.synth_labels(1, 128, 256, 19, 6, 16, 0.05, 31)
    feats = synth.synth_features(1, 96, 128, 256, [4, 8], 32)
    mod = mscs_b200.DenseContrastiveLossV2_ms(cfg)
    torch.manual_seed(9)
    gen = MT19937.from_torch_state(torch.get_rng_state().numpy().tobytes())
    with torch.no_grad():
        loss = mod(labels.to(dev), [f.to(dev) for f in feats])
    assert not loss.requires_grad
    want = loss_fp64.ms_cs_loss(labels.numpy(), [f.numpy() for f in feats], oracle_cfg(cfg, 20), gen, need_grad=False)
    assert abs(float(loss) - want["total"]) < 1e-3 * abs(want["total"])
    for s, smp in enumerate(mod.last_samples):
        assert np.array_equal(smp.idx_ref.cpu().numpy(), want["samples"][s]["idx"])


def test_single_scale_class_cross_scale_tuple(dev):
    """DenseContrastiveLossV2 with cross_scale_contrast=True returns the reference 4-tuple (V2.py:60-61)."""
    import mscs_b200
    from mscs_b200 import synth
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, cross_scale_contrast=True,
               min_views_per_class=5, max_views_per_class=30)
    labels = synth.synth_labels(2, 64, 128, 19, 5, 8, 0.05, 41)
    feat = synth.synth_features(2, 32, 64, 128, [4], 42)[0].to(dev).requires_grad_(True)
    mod = mscs_b200.DenseContrastiveLossV2(cfg)
    torch.manual_seed(3)
    loss, sampled, slabels, flag = mod(label=labels.to(dev), features=feat)
    smp = mod.last_samples[0]
    assert flag is False and sampled.shape == (smp.T, 32, smp.V) and slabels.shape == (smp.T,)
    idx, pairs = smp.idx_ref.cpu().long(), smp.pair_ref.cpu().long()
    want = feat.detach().cpu().reshape(2, 32, -1)[pairs[:, 0][:, None], :, idx].permute(0, 2, 1)
    assert torch.equal(sampled.detach().cpu(), want)
    assert torch.equal(slabels.cpu(), pairs[:, 1].float())
    loss.backward()
    assert feat.grad is not None and int((feat.grad != 0).any(dim=1).sum()) == smp.N


def test_second_backward_and_half_inputs(dev):
    """retain_graph: the second backward must not reuse the consumed pre-zeroed buffers; fp16 inputs
    (autocast-style) are accepted and the gradient comes back in the input dtype."""
    import mscs_b200
    from mscs_b200 import synth
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, max_views_per_class=20)
    labels = synth.synth_labels(2, 64, 128, 19, 5, 8, 0.05, 51).to(dev)
    feat = synth.synth_features(2, 32, 64, 128, [4], 52)[0].to(dev).requires_grad_(True)
    mod = mscs_b200.DenseContrastiveLossV2(cfg)
    torch.manual_seed(1)
    loss = mod(labels, feat)
    loss.backward(retain_graph=True)
    g1 = feat.grad.clone()
    feat.grad = None
    loss.backward()
    assert torch.allclose(feat.grad, g1, rtol=1e-4, atol=1e-9)
    fh = feat.detach().half().requires_grad_(True)
    torch.manual_seed(1)
    mod(labels, fh).backward()
    assert fh.grad.dtype == torch.float16 and torch.isfinite(fh.grad).all()


@pytest.mark.gpu
def test_fetch_logged_single_copy_and_nan_flag(dev):
    """the scalars the reference logger reads (LoggingManager.py:179-196) come back in one
    copy and agree with the per-tensor values; the device-side inf/NaN flag replaces has_inf_or_nan()."""
    import mscs_b200
    from mscs_b200 import synth
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, scales=3, weights=[1.0, 0.7, 0.4],
               cross_scale_contrast=True, min_views_per_class=5, max_views_per_class=40, max_features_total=600)
    labels = synth.synth_labels(2, 128, 256, 19, 6, 16, 0.05, 1)
    feats = synth.synth_features(2, 128, 128, 256, [4, 8, 16], 2)
    mod = mscs_b200.DenseContrastiveLossV2_ms(cfg)
    torch.manual_seed(0)
    loss = mod(labels.to(dev), [f.to(dev) for f in feats])
    got = mod.fetch_logged()
    assert got["total"] == float(loss)
    assert got["ms_losses"] == [float(x) for x in mod.ms_losses]
    assert got["cs_losses"] == [float(x) for x in mod.cs_losses]
    assert got["has_inf_or_nan"] is False and float(mod.nan_flag) == 0.0
    # the reference total: sum_s w_s ms_s + w_high_low cs(0,S-1) + w_high_mid cs(0,S-2)   (_ms.py:54,72,79)
    want = sum(w * l for w, l in zip(cfg["weights"], got["ms_losses"])) + sum(got["cs_losses"])
    assert abs(want - got["total"]) < 1e-5 * abs(want)
    bad = [f.clone() for f in feats]
    bad[1][0] = float("nan")          # every pixel of image 0 at scale 1: some of them are sampled
    torch.manual_seed(0)
    loss = mod(labels.to(dev), [f.to(dev) for f in bad])
    got = mod.fetch_logged()
    assert got["has_inf_or_nan"] is True and float(mod.nan_flag) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg4_large"])
def test_tensor_core_path_vs_fp32_validation_kernels_full_size(name, dev):
    """BASELINE.json configs 3 and 4 at full size (UPerNet/ADE20K: 151 classes, 4 scales + cross-scale terms;
    DeepLabv3/CaDIS single scale with the default and the 32768-anchor budget), where the CPU reference needs
    minutes and up to 50 GB: the tcgen05 path (bf16 operands) against the fp32 CUDA-core validation kernels on the
    SAME sampled rows -- per-term losses within 1e-3 relative, d loss / d unit rows cosine >= 0.999 (north_star
    tolerances).  The sampled indices of these configs are pinned bit-exactly by test_sampling_hashes_big."""
    import mscs_b200
    from mscs_b200 import _lib, synth
    cfg = synth.CONFIGS[name]
    labels, feats = synth.make_inputs(name)
    cls = mscs_b200.DenseContrastiveLossV2 if cfg["single_scale"] else mscs_b200.DenseContrastiveLossV2_ms
    mod = cls(dict(cfg["loss"]))
    fg = [f.to(dev).requires_grad_(True) for f in feats]
    torch.manual_seed(0)
    loss = mod(labels.to(dev), fg[0] if cfg["single_scale"] else fg)
    loss.backward()
    state = mod.last_state
    sp, job = state.sp, state.job
    assert sp.C == sp.C_pad
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    S = sp.S
    g = torch.ones((1,), dtype=torch.float32, device=dev)

    def fresh_dF():
        d = torch.zeros(sp.dF_n, dtype=torch.float32, device=dev)
        ptrs = [0] * _lib.MAX_SCALES
        lds = (C.c_int32 * _lib.MAX_SCALES)()
        for s in range(S):
            ptrs[s], lds[s] = d.data_ptr() + 4 * sp.dF_off[s], sp.C_pad
        return d, ptrs, lds

    dF_tc, ptrs, lds = fresh_dF()
    _lib.check(lib.mscs_sim_backward(C.byref(job), g.data_ptr(), _lib.ptr_array(ptrs), lds, st), "tc bwd")
    torch.cuda.synchronize()
    terms_tc = state.term_loss.clone()
    nt = state.term_loss.numel()
    total_tc = float(state.scalars[nt])       # [term losses..., total, inf/NaN flag]
    # fp32 validation kernels on the same job (fresh statistics)
    state.stats.zero_()
    fp = [0] * _lib.MAX_SCALES
    for s in range(S):
        fp[s] = state.fslab.data_ptr() + 4 * sp.foff[s][0]
    dF_v, ptrs_v, lds_v = fresh_dF()
    _lib.check(lib.mscs_debug_sim_forward_simt(C.byref(job), _lib.ptr_array(fp), st), "simt fwd")
    _lib.check(lib.mscs_debug_sim_backward_simt(C.byref(job), _lib.ptr_array(fp), g.data_ptr(), _lib.ptr_array(ptrs_v),
                                                lds_v, st), "simt bwd")
    torch.cuda.synchronize()
    terms_v = state.term_loss
    assert torch.isfinite(terms_v).all()
    rel = ((terms_tc - terms_v).abs() / terms_v.abs()).max()
    total_v = float(state.scalars[nt])
    print(f"{name}: total {total_tc:.6f} vs fp32 {total_v:.6f}; worst per-term relative difference {float(rel):.2e}")
    assert float(rel) < 1e-3 and abs(total_tc - total_v) < 1e-3 * abs(total_v)
    for s in range(S):
        N = state.samples[s].N
        a = dF_tc[sp.dF_off[s]:sp.dF_off[s] + N * sp.C_pad].double()
        b = dF_v[sp.dF_off[s]:sp.dF_off[s] + N * sp.C_pad].double()
        cs = float((a @ b) / (a.norm() * b.norm()))
        print(f"{name} set {s}: N {N} dF cosine {cs:.8f} max-abs {float((a - b).abs().max()):.3e}")
        assert cs >= 0.999


@pytest.mark.gpu
def test_generator_state_after_call_matches_host_advance(dev):
    """Q7: the call must leave the torch CPU generator where the reference's randperm calls would (n-1 draws per
    kept pair).  The product path reads the new state back from the device stream buffer; it has to equal the host
    recurrence (mscs_mt19937_advance_host, itself pinned against the reference by the golden fixtures), on a fresh
    stream (first call) and on the prefetched one (second call)."""
    import mscs_b200
    from mscs_b200 import _ops, synth
    cfg = synth.CONFIGS["cfg2"]
    labels, feats = synth.make_inputs("cfg2")
    mod = mscs_b200.DenseContrastiveLossV2_ms(dict(cfg["loss"]))
    lab, fts = labels.to(dev), [f.to(dev) for f in feats]
    torch.manual_seed(123)
    for _ in range(2):
        mt0, pos0 = _ops.torch_mt_state()
        with torch.no_grad():
            mod(lab, fts)
        got_mt, got_pos = _ops.torch_mt_state()
        from mscs_b200 import _lib
        lib = _lib.load()
        # draws = sum over kept pairs of (count - 1): count = pixels of the class in the image at that scale
        draws = 0
        for s, smp in enumerate(mod.last_samples):
            h, w = feats[s].shape[2:]
            lab_s = torch.nn.functional.interpolate(labels[:, None].float(), (h, w), mode="nearest")[:, 0].long()
            for b, c in smp.pair_ref.cpu().tolist():
                draws += int((lab_s[b] == c).sum()) - 1
        want = mt0.copy()
        cpos = C.c_int(pos0)
        _lib.check(lib.mscs_mt19937_advance_host(want.ctypes.data_as(C.c_void_p), C.byref(cpos), C.c_uint64(draws)),
                   "advance")
        assert got_pos == cpos.value
        assert np.array_equal(got_mt, want)


# ---- channels-last feature maps (torch.channels_last projector output): row gather / scatter ----------------
def _nhwc(t):
    """Same values and shape, memory order [n][h][w][C] (never plain-contiguous, even for degenerate shapes)."""
    n, Cc, h, w = t.shape
    out = torch.empty((n, h, w, Cc), dtype=t.dtype, device=t.device).permute(0, 3, 1, 2)
    out.copy_(t)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL)
def test_channels_last_small_vs_reference(name, golden, dev):
    """Channels-last inputs take the row kernels (k_gather_rows_nhwc / k_scatter_rows_nhwc); the results are held to
    the same bar against the reference's recorded outputs as the NCHW path, odd plane sizes included."""
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    mod = make_module(meta)
    fg = [_nhwc(f.to(dev)).requires_grad_(True) for f in feats]
    assert all(not f.is_contiguous() for f in fg)
    torch.set_rng_state(torch.from_numpy(z["rng_state0"]))
    loss = mod(labels.to(dev), fg[0] if meta["single_scale"] else fg)
    loss.backward()
    assert mod.last_state.sp.nhwc, "channels-last inputs must select the row kernels"
    assert abs(float(loss) - meta["total"]) < 1e-3 * abs(meta["total"])
    for s, smp in enumerate(mod.last_samples):
        assert np.array_equal(smp.idx_ref.cpu().numpy(), z[f"idx{s}"])
    for s, f in enumerate(fg):
        assert f.grad.shape == f.shape
        got, want = f.grad.cpu().numpy(), z[f"grad{s}"]
        cs, err = cosine(got, want), np.abs(got - want).max()
        print(f"{name} nhwc scale {s}: grad cosine {cs:.7f} max-abs {err:.3e}")
        assert cs >= 0.999
        assert np.array_equal(got != 0, want != 0) or np.abs(got[(got != 0) != (want != 0)]).max() < 1e-12
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


@pytest.mark.gpu
def test_channels_last_matches_nchw_full_size(dev):
    """cfg-2 at full size: the same inputs in both memory orders give the same sampled pixels, the same loss and the
    same dense gradients (the similarity kernels see identical operand rows up to the rounding of the norm)."""
    import mscs_b200
    from mscs_b200 import synth
    labels, feats = synth.make_inputs("cfg2")
    mod = mscs_b200.DenseContrastiveLossV2_ms(dict(synth.CONFIGS["cfg2"]["loss"]))
    lab = labels.to(dev)
    res = {}
    for layout in ("nchw", "nhwc"):
        fg = [(f.to(dev) if layout == "nchw" else _nhwc(f.to(dev))).requires_grad_(True) for f in feats]
        torch.manual_seed(0)
        loss = mod(lab, fg)
        loss.backward()
        assert mod.last_state.sp.nhwc == (layout == "nhwc")
        res[layout] = (float(loss), [s.pix.clone() for s in mod.last_samples], [f.grad.contiguous() for f in fg])
    (la, pa, ga), (lb, pb, gb) = res["nchw"], res["nhwc"]
    assert all(torch.equal(a, b) for a, b in zip(pa, pb))
    assert abs(la - lb) <= 2e-6 * abs(la)
    for s, (a, b) in enumerate(zip(ga, gb)):
        err, ref = float((a - b).abs().max()), float(a.abs().max())
        print(f"cfg2 nhwc vs nchw scale {s}: max-abs {err:.3e} (max {ref:.3e})")
        # the two gathers sum the squares of a row in different orders: a last-ulp difference of the norm flips the
        # bf16 rounding of a few operand elements (2^-9 relative each), and the backward accumulates with atomics in
        # an order that changes from run to run -- observed 0.6e-4 ... 1.1e-4 of the largest gradient entry
        assert err <= 3e-4 * ref
        assert torch.equal(a != 0, b != 0)


@pytest.mark.gpu
def test_channels_last_second_backward_half_and_no_grad(dev):
    import mscs_b200
    from mscs_b200 import synth
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, max_views_per_class=20)
    labels = synth.synth_labels(2, 64, 128, 19, 5, 8, 0.05, 51).to(dev)
    base = synth.synth_features(2, 32, 64, 128, [4], 52)[0].to(dev)
    feat = _nhwc(base).requires_grad_(True)
    mod = mscs_b200.DenseContrastiveLossV2(cfg)
    torch.manual_seed(1)
    loss = mod(labels, feat)
    loss.backward(retain_graph=True)
    g1 = feat.grad.clone()
    feat.grad = None
    loss.backward()
    assert torch.allclose(feat.grad, g1, rtol=1e-4, atol=1e-9)
    ref = base.clone().requires_grad_(True)
    torch.manual_seed(1)
    mod(labels, ref).backward()
    assert torch.allclose(ref.grad, g1, rtol=1e-4, atol=1e-7)
    fh = _nhwc(base.half()).requires_grad_(True)
    torch.manual_seed(1)
    mod(labels, fh).backward()
    assert fh.grad.dtype == torch.float16 and torch.isfinite(fh.grad).all() and fh.grad.shape == fh.shape
    torch.manual_seed(1)
    with torch.no_grad():
        l2 = mod(labels, _nhwc(base))
    assert abs(float(l2) - float(loss)) <= 1e-6 * abs(float(loss))


# ---- opt-in counter-based sampler ----------------------------------------------------
@pytest.mark.gpu
def test_philox_stream_kernel_matches_oracle(dev):
    """mscs_philox_stream stores the words through the inverse MT19937 tempering (the selection kernel tempers what
    it reads): tempering the buffer must give exactly the oracle's Philox4x32-10 stream."""
    import mscs_b200
    from oracle.mt19937 import temper
    from oracle.philox import stream_words
    lib = mscs_b200.load()
    seed, call, n = 0xDEADBEEF12345678, 41, 100003
    buf = torch.zeros((n + 3) // 4 * 4, dtype=torch.int32, device=dev)
    rc = lib.mscs_philox_stream(seed, call, n, buf.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.mscs_last_error()
    got = temper(buf.cpu().numpy().view(np.uint32)[:n])
    assert np.array_equal(got, stream_words(seed, call, 0, n))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_philox_sampler_matches_oracle_fed_the_same_stream(layout, dev):
    """sampler='philox': the sampled indices equal the sampling oracle's when it consumes the same counter-based stream
    (scale after scale, pair after pair, like the reference consumes its generator); the torch CPU generator is left
    alone; consecutive calls use consecutive call indices."""
    import mscs_b200
    from mscs_b200 import synth
    from oracle import sampling
    from oracle.philox import PhiloxStream
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, scales=3, weights=[1.0, 0.7, 0.4],
               cross_scale_contrast=True, min_views_per_class=5, max_views_per_class=40, max_features_total=600,
               sampler="philox", sampler_seed=99)
    labels = synth.synth_labels(2, 128, 256, 19, 6, 16, 0.05, 1)
    feats = synth.synth_features(2, 128, 128, 256, [4, 8, 16], 2)
    mod = mscs_b200.DenseContrastiveLossV2_ms(cfg)
    torch.manual_seed(5)
    rng0 = torch.get_rng_state().clone()
    seen = []
    for call in range(2):
        fg = [(f.to(dev) if layout == "nchw" else _nhwc(f.to(dev))).requires_grad_(True) for f in feats]
        loss = mod(labels.to(dev), fg)
        loss.backward()
        assert torch.isfinite(loss) and all(torch.isfinite(f.grad).all() for f in fg)
        gen = PhiloxStream(99, call)
        for s, smp in enumerate(mod.last_samples):
            want = sampling.sample_indices(labels.numpy(), feats[s].shape[-1], 20, 5, 40, 600, gen)
            assert (smp.T, smp.V) == (want["T"], want["V"])
            assert np.array_equal(smp.idx_ref.cpu().numpy(), want["idx"]), (call, s)
            assert np.array_equal(smp.pair_ref.cpu().numpy(), want["pairs"])
        seen.append(mod.last_samples[0].idx_ref.clone())
    assert not torch.equal(seen[0], seen[1])
    assert torch.equal(torch.get_rng_state(), rng0), "the counter-based sampler must not touch the torch generator"
    with pytest.raises(ValueError):
        mscs_b200.DenseContrastiveLossV2(dict(dataset="CITYSCAPES", experiment=1, sampler="nope"))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_feature_plane_larger_than_downsampled_label(layout, dev):
    """n > 1 and dl_h*dl_w < fh*fw (the reference only looks at widths, V2.py:46: a 256x512 label with 33x64 feature
    maps gives a 32x64 down-sampled label): the flat index y*dl_w + x addresses the flattened FEATURE plane of image b
    (features.view(n, c, -1)[b, :, idx], V2.py:97,123), so the image base is b*fh*fw, not b*dl_h*dl_w."""
    import mscs_b200
    from mscs_b200 import synth
    from oracle import loss_fp64
    from oracle.config import oracle_cfg
    from oracle.mt19937 import MT19937
    cfg = dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, scales=2, weights=[1.0, 0.5],
               cross_scale_contrast=True, min_views_per_class=5, max_views_per_class=25, max_features_total=2000)
    labels = synth.synth_labels(3, 256, 512, 19, 6, 16, 0.05, 21)
    g = torch.Generator().manual_seed(22)
    feats = [torch.randn(3, 32, 33, 64, generator=g), torch.randn(3, 32, 17, 32, generator=g)]
    torch.manual_seed(5)
    gen = MT19937.from_torch_state(torch.get_rng_state().numpy().tobytes())
    want = loss_fp64.ms_cs_loss(labels.numpy(), [f.numpy() for f in feats], oracle_cfg(cfg, 20), gen)
    mod = mscs_b200.DenseContrastiveLossV2_ms(cfg)
    fg = [(_nhwc(f.to(dev)) if layout == "nhwc" else f.to(dev)).requires_grad_(True) for f in feats]
    loss = mod(labels.to(dev), fg)
    loss.backward()
    assert mod.last_state.sp.nhwc == (layout == "nhwc")
    for s, smp in enumerate(mod.last_samples):
        assert (smp.dl_h * smp.dl_w) < feats[s].shape[2] * feats[s].shape[3]
        assert np.array_equal(smp.idx_ref.cpu().numpy(), want["samples"][s]["idx"])
    assert abs(float(loss) - want["total"]) < 1e-3 * abs(want["total"])
    for s, f in enumerate(fg):
        got = f.grad.cpu().numpy().astype(np.float64)
        assert np.array_equal(np.abs(got).sum(1) != 0, np.abs(want["grads"][s]).sum(1) != 0), "different pixels touched"
        assert cosine(got, want["grads"][s]) >= 0.999


def test_first_backward_of_a_fresh_process():
    """The backward is ONE C call that starts with a driver-API call (tensor-map encode) on autograd's device thread; in
    a fresh process nothing has bound a CUDA context to that thread yet (found with tools/pooled_check.py, round 2)."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, mscs_b200\n"
        "from mscs_b200 import synth\n"
        "dev = torch.device('cuda:0')\n"
        "cfg = dict(dataset='CITYSCAPES', experiment=1, temperature=0.1, scales=2, weights=[1.0, 0.5],\n"
        "           cross_scale_contrast=True, w_high_low=0.5, min_views_per_class=5, max_views_per_class=20,\n"
        "           max_features_total=2000)\n"
        "labels = synth.synth_labels(2, 64, 128, 19, 5, 8, 0.05, 3).to(dev)\n"
        "feats = [f.to(dev).requires_grad_(True) for f in synth.synth_features(2, 32, 64, 128, [4, 8], 4)]\n"
        "torch.manual_seed(0)\n"
        "loss = mscs_b200.DenseContrastiveLossV2_ms(cfg)(labels, feats)\n"
        "loss.backward()\n"
        "torch.cuda.synchronize()\n"
        "assert all(torch.isfinite(f.grad).all() and f.grad.abs().sum() > 0 for f in feats)\n"
        "print('fresh-process backward ok', float(loss))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fresh-process backward ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
