"""Projector tail evaluated on the SAMPLED rows only.

The reference's projector (models/Projector.py:49-72) ends in ``nn.Conv2d(c_prev, d, kernel_size=1)``: a d x c_prev
matrix applied to every pixel of every scale (HRNet.py:642, UPerNet.py:241), although the contrastive loss reads at most
``max_features_total`` pixels per scale -- at the HRNet-W48 Cityscapes shape 32 604 of 522 240 pixels (6 %).  This module
takes the feature maps in front of that last convolution and evaluates

    DenseContrastiveLossV2_ms(config)(label, [conv_s(z_s) for s])          (losses/DenseContrastiveLossV2_ms.py:44-82)

without ever forming the dense projector output or its dense gradient:

    K1 sampling on the label map (unchanged, bit-exact)          -> pixel -> row maps
    raw gather of the sampled c_prev-vectors (mscs_gather_rows_raw)   Z_s  (N_s, c_prev)
    P_s = Z_s W_s^T + b_s                                          plain library GEMM (N_s x c_prev x d)
    row normalisation -> bf16 operands (mscs_gather_rows_nhwc_batch on the contiguous rows)
    K3 / K4 as in the drop-in classes
    normalisation backward (mscs_scatter_rows_nhwc_batch) -> dP_s;  dZ_s = dP_s W_s, dW_s = dP_s^T Z_s, db_s = sum dP_s
    raw scatter of dZ_s into the dense (n, c_prev, h, w) gradient (mscs_scatter_rows_raw)

Same loss, same gradients w.r.t. z_s, W_s and b_s as the dense evaluation (tests/test_gpu_projector.py compares with
``nn.Conv2d`` + the reference's own loss class).  It is NOT a reference class name: the reference has no such module,
so it sits next to the drop-in classes with its own, explicit interface.
"""
import torch
import torch.nn as nn

from . import _lib, _ops
from .losses import _spec_from_config
from .datasets import class_facts


class _ProjTailFn(torch.autograd.Function):
    """(labels, spec, holder, S, z_0..z_{S-1}, W_0..W_{S-1}, b_0..b_{S-1}) -> (total 0-d, term losses)."""

    @staticmethod
    def forward(ctx, labels, spec, holder, S, *tensors):
        lib = _lib.load()
        zs, Ws, bs = tensors[:S], tensors[S:2 * S], tensors[2 * S:3 * S]
        dev = zs[0].device
        for t in tensors:
            _ops._require_device(t)
        if not lib.mscs_device_ok():
            raise RuntimeError("mscs_b200 needs a compute-capability 10.x (B200) device; no fallback exists")
        labels = labels.to(dev)
        if labels.dtype != torch.int64:
            labels = labels.long()
        labels = labels.contiguous()
        z32 = [z.detach().float().contiguous() for z in zs]
        with torch.cuda.device(dev), _ops._pin_stream():
            st = _ops._stream()
            shapes = [tuple(z.shape) for z in z32]
            for n, c, h, w in shapes:
                if (h * w) % 8 != 0:
                    raise NotImplementedError("the projector-tail path needs feature planes that are a multiple of 8 pixels")
            # K1 (host-driven order); the generator bookkeeping + prefetch of the next call's stream run after the sweeps
            samples, finish_rng = _ops.sample_anchors(labels, [(s[2], s[3]) for s in shapes], spec, defer_rng=True)
            d = Ws[0].shape[0]
            C_pad = (d + 63) // 64 * 64
            sets, slots, Zs, iotas = [], [], [], []
            items = (_lib.RowsItem * S)()
            Ps = []
            for s in range(S):
                n, c, h, w = shapes[s]
                N = samples[s].N
                slot = torch.full((n * h * w,), -1, dtype=torch.int32, device=dev)
                _lib.check(lib.mscs_slot_map(samples[s].ptr(2), N, n * h * w, slot.data_ptr(), st), "mscs_slot_map")
                Z = torch.empty((N, c), dtype=torch.float32, device=dev)
                _lib.check(lib.mscs_gather_rows_raw(z32[s].data_ptr(), n, c, h * w, slot.data_ptr(), Z.data_ptr(), st),
                           "mscs_gather_rows_raw")
                # the 1x1 convolution of the sampled pixels: a plain library GEMM (fp32, as nn.Conv2d computes it)
                P = torch.addmm(bs[s].detach().float(), Z, Ws[s].detach().float().reshape(d, c).t())
                N_pad = (N + 255) // 256 * 256
                aset = _ops.AnchorSet(N=N, C=d, C_pad=C_pad,
                                      bf16=torch.empty((N_pad, C_pad), dtype=torch.bfloat16, device=dev),
                                      f32=torch.empty((N, d), dtype=torch.float32, device=dev),
                                      inv_norm=torch.empty((N,), dtype=torch.float32, device=dev))
                iota = torch.arange(N, dtype=torch.int32, device=dev)
                it = items[s]
                it.feat, it.pix, it.n_rows_dev, it.rows, it.C = P.data_ptr(), iota.data_ptr(), None, N, d
                it.anc_bf16, it.anc_f32, it.inv_norm = aset.bf16.data_ptr(), aset.f32.data_ptr(), aset.inv_norm.data_ptr()
                sets.append(aset); slots.append(slot); Zs.append(Z); iotas.append(iota); Ps.append(P)
            _lib.check(lib.mscs_gather_rows_nhwc_batch(items, S, st), "mscs_gather_rows_nhwc_batch")
            state = _ops.build_job(spec, samples, sets, single_scale=False)
            _ops.sim_forward(state)
            finish_rng()
        holder["samples"], holder["state"] = samples, state
        ctx.state, ctx.sets, ctx.slots, ctx.Zs, ctx.iotas, ctx.shapes, ctx.S = state, sets, slots, Zs, iotas, shapes, S
        ctx.save_for_backward(*Ws)
        ctx.dtypes = [t.dtype for t in tensors]
        ctx.keep = (Ps, z32, labels)
        total, state.total = state.total, None          # the output must not be reachable from ctx (cycle through C++)
        terms = state.term_loss
        ctx.mark_non_differentiable(terms)
        ctx.set_materialize_grads(False)
        return total, terms

    @staticmethod
    def backward(ctx, grad_total, _grad_terms):
        S = ctx.S
        if grad_total is None:
            return (None,) * (4 + 3 * S)
        lib = _lib.load()
        Ws = ctx.saved_tensors
        dev = ctx.Zs[0].device
        with torch.cuda.device(dev), _ops._pin_stream():
            st = _ops._stream()
            dFs = _ops.sim_backward(ctx.state, ctx.sets, grad_total)                   # K4: d loss / d unit rows
            items = (_lib.RowsItem * S)()
            dPs = []
            for s in range(S):
                a = ctx.sets[s]
                dP = torch.empty((a.N, a.C), dtype=torch.float32, device=dev)
                it = items[s]
                it.pix, it.rows, it.C = ctx.iotas[s].data_ptr(), a.N, a.C
                it.anc_f32, it.inv_norm = a.f32.data_ptr(), a.inv_norm.data_ptr()
                it.dF, it.ldF, it.dfeat = dFs[s].data_ptr(), dFs[s].shape[1], dP.data_ptr()
                dPs.append(dP)
            _lib.check(lib.mscs_scatter_rows_nhwc_batch(items, S, st), "mscs_scatter_rows_nhwc_batch")   # normalise^T
            gz, gW, gb = [], [], []
            needs = ctx.needs_input_grad[4:]
            for s in range(S):
                n, c, h, w = ctx.shapes[s]
                W2 = Ws[s].detach().float().reshape(-1, c)
                if needs[s]:
                    dZ = dPs[s] @ W2                                                                  # (N, c_prev)
                    dz = torch.zeros((n, c, h, w), dtype=torch.float32, device=dev)
                    _lib.check(lib.mscs_scatter_rows_raw(dZ.data_ptr(), c, ctx.slots[s].data_ptr(), n, c, h * w,
                                                         dz.data_ptr(), st), "mscs_scatter_rows_raw")
                    gz.append(dz if ctx.dtypes[s] == torch.float32 else dz.to(ctx.dtypes[s]))
                else:
                    gz.append(None)
                gW.append((dPs[s].t() @ ctx.Zs[s]).reshape(Ws[s].shape).to(ctx.dtypes[S + s]) if needs[S + s] else None)
                gb.append(dPs[s].sum(0).to(ctx.dtypes[2 * S + s]) if needs[2 * S + s] else None)
        return (None, None, None, None, *gz, *gW, *gb)


class ProjectorTailContrastive_ms(nn.Module):
    """``forward(label, pre_features)`` == ``DenseContrastiveLossV2_ms(config)(label, [tail_s(z_s) for s])`` with
    ``tail_s = nn.Conv2d(c_prev_s, d, 1)`` -- the LAST layer of the reference's projector (Projector.py:69), owned by
    this module as ``self.tails[s]`` (``from_projector`` takes them out of a reference ``Projector``).

    Same config keys as DenseContrastiveLossV2_ms (losses.py); exposes ``ms_losses`` / ``cs_losses`` /
    ``cross_scale_contrast`` like it, so the reference's logger reads it the same way."""

    def __init__(self, config, c_prev, d=256):
        super().__init__()
        self.scales = config["scales"] if "scales" in config else 2
        self.weights = config["weights"] if "weights" in config else [1.0] * self.scales
        assert self.scales == len(self.weights)
        c_prev = [c_prev] * self.scales if isinstance(c_prev, int) else list(c_prev)
        assert len(c_prev) == self.scales
        self.num_all_classes = class_facts(config["dataset"], config["experiment"])[0]
        self._spec = _spec_from_config(config, self.scales, self.weights, ms=True)
        if self._spec.sampler != "reference":
            raise NotImplementedError("the projector-tail path uses the reference sampler")
        self.cross_scale_contrast = self._spec.cross_scale
        self.tails = nn.ModuleList([nn.Conv2d(c, d, kernel_size=1, stride=1) for c in c_prev])
        self.ms_losses, self.cs_losses = [], []
        self.last_samples = None

    @classmethod
    def from_projector(cls, config, projector):
        """Takes over the final 1x1 convolution of every ``project{i}`` of a reference ``Projector`` (is_ms) and
        returns ``(module, bodies)`` with ``bodies[i]`` = that Sequential without its last layer."""
        names = [f"project{i}" for i in range(len(projector.c_in))] if projector.is_ms else ["project"]
        seqs = [getattr(projector, nm) for nm in names]
        lasts = [seq[-1] for seq in seqs]
        mod = cls(config, [l.in_channels for l in lasts], lasts[0].out_channels)
        for mine, theirs in zip(mod.tails, lasts):
            mine.load_state_dict(theirs.state_dict())
        return mod, [nn.Sequential(*list(seq.children())[:-1]) for seq in seqs]

    def forward(self, label, pre_features):
        feats = list(pre_features[:self.scales])
        if len(feats) < self.scales:
            raise IndexError("list index out of range")
        holder = {}
        total, terms = _ProjTailFn.apply(label, self._spec, holder, self.scales, *feats,
                                         *[t.weight for t in self.tails], *[t.bias for t in self.tails])
        state = holder["state"]
        self.last_samples, self.last_state = holder["samples"], state
        self.ms_losses = [terms[s] for s in range(state.num_ms)]
        self.cs_losses = [terms[i] for i in state.cs_logged]
        return total
