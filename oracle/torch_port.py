"""fp32 torch restatement of the reference loss with autograd -- the timed CPU baseline ("port").
TEST INFRASTRUCTURE.

/root/reference is a Python package that cannot travel to the GPU box, so ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs time this port instead.  It executes the same ATen
op sequence as the reference on the same host threads (per-pair nonzero + CPU randperm + gather,
F.normalize, dense N x N fp32 matmul, float masks, exp/log/sum, autograd backward), so its
wall-clock is representative of the reference's; ``tests/golden/make_golden.py`` checks its
loss and sampled sets against the real reference bit-for-bit on the golden inputs.

Follows losses/DenseContrastiveLossV2.py:44-206 and losses/DenseContrastiveLossV2_ms.py:44-161.
"""
import torch
import torch.nn.functional as tnf


def _downsample(label, feat_w):
    n, H, W = label.shape
    s = int(W // feat_w)
    return tnf.interpolate(label[:, None].float(), (H // s, W // s), mode="nearest").long().view(n, -1)


def sample(label, feat, A, min_views, max_views, max_total):
    """-> sampled (T,C,V) fp32 (differentiable w.r.t. feat), class ids (T,), idx (T,V), pairs."""
    n, C = feat.shape[:2]
    dl = _downsample(label, feat.shape[-1])
    onehot = dl[:, :, None] == torch.arange(A)[None, None, :]
    counts = onehot.sum(1)
    bsel, csel = torch.where(counts[:, :-1] >= min_views)
    min_count = int(torch.min(counts[bsel, csel]))
    T = bsel.shape[0]
    V = min_count if max_views == 1 else min(min_count, max_views)
    if V * T > max_total:
        V = max_total // T
    fl = feat.view(n, C, -1)
    out = torch.zeros((T, C, V), dtype=torch.float)
    idx = torch.empty((T, V), dtype=torch.long)
    for k in range(T):
        where_c = onehot[bsel[k], :, csel[k]].nonzero().squeeze()
        take = where_c[torch.randperm(where_c.shape[0])[:V]]
        out[k] = fl[bsel[k], :, take]
        idx[k] = take
    return out, csel.float(), idx, torch.stack([bsel, csel], 1)


def _flat(feats, labels):
    f = tnf.normalize(feats, p=2, dim=1).transpose(1, 2)
    T, V, C = f.shape
    return f.contiguous().view(-1, C), labels.view(-1, 1).repeat(1, V).view(-1, 1)


def _masked_nll(logits, pos, neg, guard):
    e = torch.exp(logits)
    neg_sum = (e * neg).sum(1, keepdim=True)
    log_prob = logits - torch.log(e + neg_sum)
    p = pos.sum(1)
    div = torch.where(p > 0, p, torch.ones_like(p)) if guard else p
    return -((pos * log_prob).sum(1) / div).mean()


def single_scale_term(feats, labels, tau):
    f, y = _flat(feats, labels)
    same = torch.eq(y, y.t()).float()
    eye_off = torch.ones_like(same).scatter_(1, torch.arange(f.shape[0]).view(-1, 1), 0)
    return _masked_nll(torch.matmul(f, f.t()) / tau, same * eye_off, 1 - same, guard=False)


def cross_scale_term(f1, y1, f2, y2, tau):
    a, ya = _flat(f1, y1)
    k, yk = _flat(f2, y2)
    same = torch.eq(ya, yk.t()).float()
    return _masked_nll(torch.matmul(a, k.t()) / tau, same, 1 - same, guard=True)


def ms_cs_loss(label, feats, cfg):
    """cfg as in oracle.loss_fp64.ms_cs_loss.  Returns (total, ms list, cs list, idx list)."""
    S = len(feats)
    weights = cfg.get("weights") or [1.0] * S
    total = torch.tensor(0.0)
    sets, ms, cs, idxs = [], [], [], []
    for s in range(S):
        sf, sl, idx, _ = sample(label, feats[s], cfg["num_all_classes"], cfg["min_views"],
                                cfg["max_views"], cfg["max_total"])
        l = single_scale_term(sf, sl, cfg["temperature"])
        total = total + weights[s] * l
        sets.append((sf, sl))
        ms.append(l.detach())
        idxs.append(idx)
    if cfg.get("cross_scale") and S > 1:
        det = (lambda t: t.detach()) if cfg.get("detach_deepest") else (lambda t: t)
        l = cross_scale_term(sets[0][0], sets[0][1], det(sets[-1][0]), sets[-1][1], cfg["cs_temperature"])
        total = total + cfg.get("w_high_low", 1.0) * l
        if not cfg.get("detach_deepest"):
            cs.append(l.detach())
        if S > 2:
            l = cross_scale_term(sets[0][0], sets[0][1], det(sets[-2][0]), sets[-2][1], cfg["cs_temperature"])
            total = total + cfg.get("w_high_mid", 1.0) * l
            cs.append(l.detach())
    return total, ms, cs, idxs
