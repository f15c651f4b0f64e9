"""Chunked fp64 restatement of the ms+cs dense contrastive loss and its gradient (numpy).
TEST INFRASTRUCTURE.

Follows
  losses/DenseContrastiveLossV2.py:127-192     contrastive_loss / get_masks2 / get_loss
  losses/DenseContrastiveLossV2_ms.py:44-82    scale combination, cross-scale terms
  losses/DenseContrastiveLossV2_ms.py:84-161   cross-scale contrastive_loss / InfoNce_loss
and the analytic gradient of SURVEY.md Appendix A (checked against autograd of the executable
reference by tests/golden/make_golden.py).  Never materialises more than ``chunk x N`` logits,
so it also serves the pooled 64-image configuration where the reference cannot run.
"""
import numpy as np

from .sampling import sample_indices


def gather_normalize(feat, pairs, idx):
    """feat (n,C,h,w) -> unit rows (T*V, C) fp64 in reference order k*V+v, plus the raw norms.
    F.normalize(p=2, dim=1, eps=1e-12): x / max(||x||, eps)   (V2.py:138)."""
    n, C = feat.shape[:2]
    flat = np.asarray(feat, dtype=np.float64).reshape(n, C, -1)
    T, V = idx.shape
    X = np.empty((T * V, C), dtype=np.float64)
    for k in range(T):
        X[k * V:(k + 1) * V] = flat[pairs[k, 0]][:, idx[k]].T
    norm = np.sqrt((X * X).sum(1))
    return X / np.maximum(norm, 1e-12)[:, None], norm


def _row_pass(Fa, ya, Fk, yk, tau, self_mask, chunk):
    """Sweep A: N_i = sum over negatives of exp(l_ik)."""
    N1 = Fa.shape[0]
    neg = np.zeros(N1)
    for r0 in range(0, N1, chunk):
        r1 = min(N1, r0 + chunk)
        E = np.exp(Fa[r0:r1] @ Fk.T / tau)
        neg[r0:r1] = (E * (ya[r0:r1, None] != yk[None, :])).sum(1)
    return neg


def term(Fa, ya, Fk, yk, tau, self_mask, need_grad=True, chunk=1024):
    """One contrastive term.  rows = anchors, columns = keys.

    self_mask=True  -> single-scale term (V2.py:164-188): positives exclude i==j, divisor P_i
                       (0/0 -> NaN exactly like the reference);
    self_mask=False -> cross-scale InfoNCE (_ms.py:128-156): divisor max(P_i, 1).
    Returns (loss, dFa, dFk, stats) with d(loss)/d(unit rows); for self_mask the caller adds
    dFa + dFk (same matrix).  stats = dict(neg, possum, S, P).
    """
    N1, N2 = Fa.shape[0], Fk.shape[0]
    neg = _row_pass(Fa, ya, Fk, yk, tau, self_mask, chunk)
    possum, S, P = np.zeros(N1), np.zeros(N1), np.zeros(N1)
    for r0 in range(0, N1, chunk):
        r1 = min(N1, r0 + chunk)
        L = Fa[r0:r1] @ Fk.T / tau
        pos = ya[r0:r1, None] == yk[None, :]
        if self_mask:
            pos[np.arange(r1 - r0), np.arange(r0, r1)] = False
        den = np.exp(L) + neg[r0:r1, None]
        possum[r0:r1] = ((L - np.log(den)) * pos).sum(1)
        S[r0:r1] = (pos / den).sum(1)
        P[r0:r1] = pos.sum(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        div = P if self_mask else np.where(P > 0, P, 1.0)
        loss = float((-possum / div).mean())
    stats = dict(neg=neg, possum=possum, S=S, P=P)
    if not need_grad:
        return loss, None, None, stats
    dFa, dFk = np.zeros_like(Fa), np.zeros_like(Fk)
    with np.errstate(invalid="ignore", divide="ignore"):
        inv = 1.0 / (div * N1)
    for r0 in range(0, N1, chunk):
        r1 = min(N1, r0 + chunk)
        E = np.exp(Fa[r0:r1] @ Fk.T / tau)
        pos = ya[r0:r1, None] == yk[None, :]
        negm = ~pos
        if self_mask:
            pos[np.arange(r1 - r0), np.arange(r0, r1)] = False
        n_i = neg[r0:r1, None]
        G = np.where(pos, -inv[r0:r1, None] * n_i / (E + n_i), 0.0)
        G += np.where(negm, inv[r0:r1, None] * S[r0:r1, None] * E, 0.0)
        dFa[r0:r1] = G @ Fk / tau
        dFk += G.T @ Fa[r0:r1] / tau
    return loss, dFa, dFk, stats


def normalize_backward(dF, F, norm):
    """d/dx of x / max(||x||, eps):  (dF - f (f.dF)) / ||x||   (plain dF/eps below the clamp)."""
    small = norm <= 1e-12
    dot = (dF * F).sum(1, keepdims=True)
    dx = (dF - F * dot) / np.maximum(norm, 1e-12)[:, None]
    if small.any():
        dx[small] = dF[small] / 1e-12
    return dx


def scatter_dense(dx, pairs, idx, shape):
    """dense zero grad with the sampled columns written (autograd of V2.py:123)."""
    n, C, h, w = shape
    out = np.zeros((n, C, h * w))
    T, V = idx.shape
    for k in range(T):
        out[pairs[k, 0]][:, idx[k]] += dx[k * V:(k + 1) * V].T
    return out.reshape(shape)


def ms_cs_loss(label, feats, cfg, gen, need_grad=True, chunk=1024, samples=None, dense=True):
    """The whole path: DenseContrastiveLossV2_ms.forward (or the single-scale class when
    ``cfg['single_scale']``) + gradient w.r.t. every feature map.

    cfg keys: num_all_classes, temperature, cs_temperature, min_views, max_views, max_total,
    weights, cross_scale, detach_deepest, w_high_low, w_high_mid.
    Returns dict(total, ms, cs, samples, grads, grad_rows); ``dense=False`` skips the dense scatter (grads = None)
    and only returns ``grad_rows``: per scale the (T*V, C) gradient rows at the sampled pixels in reference order
    k*V+v -- the only non-zero part of the dense gradient (sizes where a dense fp64 map is too large to hold).
    """
    A = cfg["num_all_classes"]
    S = len(feats)
    if samples is None:
        samples = [sample_indices(label, feats[s].shape[-1], A, cfg["min_views"], cfg["max_views"],
                                  cfg["max_total"], gen) for s in range(S)]
    Fs, norms, ys = [], [], []
    for s in range(S):
        F, nr = gather_normalize(feats[s], samples[s]["pairs"], samples[s]["idx"])
        Fs.append(F)
        norms.append(nr)
        ys.append(np.repeat(samples[s]["pairs"][:, 1], samples[s]["V"]))
    dFs = [np.zeros_like(F) for F in Fs]
    weights = cfg.get("weights") or [1.0] * S
    ms, cs, total = [], [], 0.0
    for s in range(S):
        l, da, dk, _ = term(Fs[s], ys[s], Fs[s], ys[s], cfg["temperature"], True, need_grad, chunk)
        ms.append(l)
        total += weights[s] * l
        if need_grad:
            dFs[s] += weights[s] * (da + dk)
    if cfg.get("cross_scale") and S > 1:
        keys = [(S - 1, cfg.get("w_high_low", 1.0))]
        if S > 2:
            keys.append((S - 2, cfg.get("w_high_mid", 1.0)))
        for j, (ks, w) in enumerate(keys):
            l, da, dk, _ = term(Fs[0], ys[0], Fs[ks], ys[ks], cfg["cs_temperature"], False, need_grad, chunk)
            total += w * l
            # _ms.py:66-70: with detach_deepest the first cs loss is not appended to cs_losses
            if not (cfg.get("detach_deepest") and j == 0):
                cs.append(l)
            if need_grad:
                dFs[0] += w * da
                if not cfg.get("detach_deepest"):
                    dFs[ks] += w * dk
    grads = rows = None
    if need_grad:
        rows = [normalize_backward(dFs[s], Fs[s], norms[s]) for s in range(S)]
        if dense:
            grads = [scatter_dense(rows[s], samples[s]["pairs"], samples[s]["idx"], feats[s].shape) for s in range(S)]
    return dict(total=total, ms=ms, cs=cs, samples=samples, grads=grads, grad_rows=rows)
