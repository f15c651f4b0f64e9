"""Anchor sampling of the reference, restated in numpy.  TEST INFRASTRUCTURE.

Follows ``losses/DenseContrastiveLossV2.py``:
  :194-206 get_dist_and_classes   nearest down-sampling of the label map (float32 round trip)
  :86-125  sample_anchors_fast    class histogram, pair list, V rule, per-pair randperm
  :64-84   _select_views_per_class
"""
import numpy as np

from .mt19937 import randperm_prefix


def nearest_source_index(out_size, in_size):
    """ATen ``upsample_nearest`` source index (third-party, torch 2.11): identity when sizes
    match, otherwise ``min(int(floorf(dst * float(in)/out)), in-1)`` in float32."""
    if out_size == in_size:
        return np.arange(out_size, dtype=np.int64)
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


def downsample_labels(label, feat_w):
    """V2.py:46 + :202-206 -> (n, H//s, W//s) int64 with s = W // feat_w."""
    label = np.asarray(label)
    n, H, W = label.shape
    s = int(W // feat_w)
    oh, ow = H // s, W // s
    ys, xs = nearest_source_index(oh, H), nearest_source_index(ow, W)
    dl = label[:, ys][:, :, xs]
    return dl.astype(np.float32).astype(np.int64)       # the reference's float round trip


def views_per_class(min_count, T, max_views, max_total):
    """V2.py:64-84.  Returns (V, log_this_step)."""
    log = False
    if max_views == 1:
        V = int(min_count)
    else:
        V = min(int(min_count), max_views)
        log = V == max_views
    if V * T > max_total:
        V = max_total // T
        log = True
    return V, log


def sample_plan(dl, A, min_views):
    """Histogram + pair list (V2.py:101-111).  Returns counts (n,A), pairs [(b,c)...] row-major."""
    n = dl.shape[0]
    flat = dl.reshape(n, -1)
    counts = np.stack([(flat == c).sum(1) for c in range(A)], axis=1).astype(np.int64)
    bs, cs = np.nonzero(counts[:, :-1] >= min_views)      # torch.where order: row-major (b, c)
    return counts, list(zip(bs.tolist(), cs.tolist()))


def sample_indices(label, feat_w, A, min_views, max_views, max_total, gen):
    """Full sampling for one scale.

    Returns dict(T, V, pairs (T,2), counts (n,A), idx (T,V) flat positions y*w+x in the
    down-sampled map, log_this_step).  ``gen`` is an ``MT19937`` positioned where the torch CPU
    default generator is when the reference is called; it is advanced exactly as torch is.
    """
    dl = downsample_labels(label, feat_w)
    counts, pairs = sample_plan(dl, A, min_views)
    if not pairs:
        raise RuntimeError("no (image, class) pair with >= min_views_per_class pixels "
                           "(reference: torch.min of an empty tensor raises, V2.py:110)")
    T = len(pairs)
    min_count = min(int(counts[b, c]) for b, c in pairs)
    V, log = views_per_class(min_count, T, max_views, max_total)
    idx = np.empty((T, V), dtype=np.int64)
    flat = dl.reshape(dl.shape[0], -1)
    for k, (b, c) in enumerate(pairs):
        pos = np.flatnonzero(flat[b] == c)
        if pos.shape[0] == 1:
            raise IndexError("class with a single pixel: reference fails on 0-d squeeze (V2.py:119-121)")
        perm = randperm_prefix(pos.shape[0], V, gen)
        idx[k] = pos[perm]
    return dict(T=T, V=V, pairs=np.asarray(pairs, dtype=np.int64), counts=counts, idx=idx,
                log_this_step=log, dl_shape=dl.shape[1:])
