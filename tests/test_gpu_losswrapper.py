"""GPU tests through the reference's own boundary: the UNMODIFIED ``losses/LossWrapper.py`` (staged under
``oracle/_ref/`` by oracle/make_ref.py -- the GPU box has no /root/reference) builds this repository's classes by
name (LossWrapper.py:33), calls ``module(labels, deep_features)`` (:68-71), multiplies the result IN PLACE by the
loss weight (:90), detaches it for the logger (:91), copies ``ms_losses`` / ``cs_losses`` into ``loss_vals``
(:94-101) and adds it to ``total_loss`` (:102); the test then calls ``total_loss.backward()`` like
``forward_step`` does.  Checked against (a) the recorded outputs of the reference (tests/golden), and (b) the
reference's OWN loss classes run live on the same GPU through the same wrapper (fp32 ATen: torch's CUDA matmul does
not use TF32 unless asked to), at every BASELINE configuration whose N x N temporaries fit the device.

Tolerances (north_star): loss <= 1e-3 relative, gradient cosine >= 0.999, max-abs printed.
"""
import numpy as np
import pytest
import torch

from helpers import cosine, load_npz, small_case_inputs

pytestmark = pytest.mark.gpu

MS, SS = "DenseContrastiveLossV2_ms", "DenseContrastiveLossV2"


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    if ref_loader.find_root() is None:
        pytest.fail("oracle/_ref/ is missing: __graft_entry__.build() stages it in the build container")
    return ref_loader.load(cpu=False)


class _installed:
    """Our classes visible under the reference's names in the reference modules (losses.install_into_reference),
    the reference's own classes put back afterwards."""

    def __init__(self, ref):
        self.ref = ref

    def __enter__(self):
        import mscs_b200
        mscs_b200.install_into_reference()
        return mscs_b200

    def __exit__(self, *exc):
        for m in (self.ref.losses, self.ref.wrapper_module):
            m.DenseContrastiveLossV2 = self.ref.DenseContrastiveLossV2
            m.DenseContrastiveLossV2_ms = self.ref.DenseContrastiveLossV2_ms


def _wrapper_cfg(loss_cfg, key, weight, dev):
    cfg = dict(loss_cfg)
    cfg.update(losses={key: weight}, device=dev)
    return cfg


def _run_wrapper(ref, cfg, key, labels, feats, rng_state, dev):
    """LossWrapper(cfg) -> forward(None, labels, deep_features=...) -> total_loss.backward(), as forward_step does."""
    lw = ref.LossWrapper(cfg)
    fg = [f.detach().clone().to(dev).requires_grad_(True) for f in feats]
    torch.set_rng_state(rng_state)
    total = lw(None, labels.to(dev), deep_features=fg[0] if key == SS else fg)
    total.backward()
    torch.cuda.synchronize()
    return lw, total, fg


@pytest.mark.parametrize("name", ["tiny_ms", "tiny_ms_detach", "tiny_ss", "odd_ss"])
def test_unmodified_losswrapper_small_vs_recorded_reference(name, golden, ref, dev):
    meta = golden[name]
    labels, feats, z = small_case_inputs(name)
    key, w = (SS if meta["single_scale"] else MS), 0.1
    with _installed(ref) as mscs_b200:
        cfg = _wrapper_cfg(meta["loss_cfg"], key, w, dev)
        lw, total, fg = _run_wrapper(ref, cfg, key, labels, feats, torch.from_numpy(z["rng_state0"]), dev)
        assert type(lw.loss_classes[key]) is getattr(mscs_b200, key)
    assert total.dim() == 0 and total.dtype == torch.float32
    assert abs(float(total) - w * meta["total"]) < 1e-3 * abs(w * meta["total"])
    assert abs(float(lw.loss_vals[key]) - w * meta["total"]) < 1e-3 * abs(w * meta["total"])
    assert not lw.loss_vals[key].requires_grad
    if key == MS:
        # logged per-scale values are the UNWEIGHTED term losses (LossWrapper.py:96-101)
        for i, v in enumerate(meta["ms"]):
            assert abs(float(lw.loss_vals[f"{MS}_ms{i}"]) - v) < 1e-3 * abs(v)
        for i, v in enumerate(meta["cs"]):
            assert abs(float(lw.loss_vals[f"{MS}_cs{i}"]) - v) < 1e-3 * abs(v)
        assert f"{MS}_ms{len(meta['ms'])}" not in lw.loss_vals and f"{MS}_cs{len(meta['cs'])}" not in lw.loss_vals
        # the in-place multiply of the wrapper must not leak into the logged scalars of the module
        logged = lw.loss_classes[key].fetch_logged()
        assert abs(logged["total"] - meta["total"]) < 1e-3 * abs(meta["total"])
    for s, f in enumerate(fg):
        got, want = f.grad.cpu().numpy(), w * z[f"grad{s}"]
        cs, err = cosine(got, want), np.abs(got - want).max()
        print(f"{name} via LossWrapper, scale {s}: grad cosine {cs:.7f} max-abs {err:.3e} (max {np.abs(want).max():.3e})")
        assert cs >= 0.999
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


def test_unmodified_losswrapper_cfg2_vs_recorded_reference(golden, ref, dev):
    from mscs_b200 import synth
    meta, z = golden["cfg2"], load_npz("cfg2")
    labels, feats = synth.make_inputs("cfg2")
    w = 0.1
    with _installed(ref):
        lw, total, fg = _run_wrapper(ref, _wrapper_cfg(meta["loss_cfg"], MS, w, dev), MS, labels, feats,
                                     torch.from_numpy(z["rng_state0"]), dev)
    assert abs(float(total) - w * meta["total"]) < 1e-3 * abs(w * meta["total"])
    for i, v in enumerate(meta["ms"]):
        assert abs(float(lw.loss_vals[f"{MS}_ms{i}"]) - v) < 1e-3 * abs(v)
    for i, v in enumerate(meta["cs"]):
        assert abs(float(lw.loss_vals[f"{MS}_cs{i}"]) - v) < 1e-3 * abs(v)
    for s, f in enumerate(fg):
        tot = float(f.grad.double().norm())
        assert abs(tot - w * meta["grad_l2"][s]) < 1e-2 * w * meta["grad_l2"][s]
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(z["rng_state1"]))


LIVE = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg4_large"]


@pytest.mark.parametrize("name", LIVE)
def test_live_reference_on_gpu_same_wrapper_same_inputs(name, ref, dev):
    """The reference's own classes and ours behind the SAME unmodified LossWrapper, same inputs, same generator state,
    both on this GPU: total loss, every logged scalar and the full dense gradients."""
    from mscs_b200 import synth
    cfgd = synth.CONFIGS[name]
    labels, feats = synth.make_inputs(name)
    key, w = (SS if cfgd["single_scale"] else MS), 0.1
    torch.manual_seed(0)
    state0 = torch.get_rng_state()
    cfg = _wrapper_cfg(cfgd["loss"], key, w, dev)
    free, _total = torch.cuda.mem_get_info()
    n_big = 32768 if name == "cfg4_large" else 10000
    if free < 14 * 4 * n_big * n_big + (4 << 30):
        pytest.skip(f"{name}: the reference's N x N fp32 temporaries do not fit ({free >> 30} GiB free)")
    with _installed(ref):
        lw_o, total_o, fg_o = _run_wrapper(ref, cfg, key, labels, feats, state0, dev)
        state_o = torch.get_rng_state()
        vals_o = {k: float(v) for k, v in lw_o.loss_vals.items()}
        grads_o = [f.grad.clone() for f in fg_o]
        del lw_o, fg_o
    torch.cuda.empty_cache()
    with torch.cuda.device(dev):          # the reference's hard-coded .cuda() calls use the current device (Q9)
        lw_r, total_r, fg_r = _run_wrapper(ref, cfg, key, labels, feats, state0, dev)
    assert type(lw_r.loss_classes[key]) is getattr(ref, key)
    assert torch.equal(torch.get_rng_state(), state_o), "generator state after the call differs from the reference's"
    vals_r = {k: float(v) for k, v in lw_r.loss_vals.items()}
    assert set(vals_r) == set(vals_o), (sorted(vals_r), sorted(vals_o))
    for k, v in vals_r.items():
        rel = abs(vals_o[k] - v) / abs(v)
        print(f"{name} {k}: ours {vals_o[k]:.7f} reference {v:.7f} rel {rel:.2e}")
        assert rel < 1e-3, (k, vals_o[k], v)
    assert abs(float(total_o) - float(total_r)) < 1e-3 * abs(float(total_r))
    for s, (go, fr) in enumerate(zip(grads_o, fg_r)):
        gr = fr.grad
        # the same PIXELS carry a gradient (single channels of a sampled pixel may be exactly 0 on one side only)
        assert torch.equal((go != 0).any(dim=1), (gr != 0).any(dim=1)), "different pixels carry a gradient"
        a, b = go.double().flatten(), gr.double().flatten()
        cs = float((a @ b) / (a.norm() * b.norm()))
        err, mx = float((a - b).abs().max()), float(b.abs().max())
        print(f"{name} scale {s}: dense grad cosine {cs:.8f} max-abs {err:.3e} (reference max {mx:.3e})")
        assert cs >= 0.999
        assert err < 0.05 * mx
    del lw_r, fg_r
    torch.cuda.empty_cache()
