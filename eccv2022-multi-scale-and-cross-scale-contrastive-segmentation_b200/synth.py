"""Deterministic synthetic inputs for the benchmark / parity configurations.

Everything is drawn from torch CPU generators with fixed seeds so that the inputs are
identical here, on the GPU box and in the golden-vector generator.
"""
import torch


def synth_labels(n, H, W, num_real, k, cell, ignore_frac, seed):
    """Blocky label maps: every ``cell x cell`` block carries one of ``k`` classes of image b
    (or the ignore id ``num_real`` with probability ``ignore_frac``).  int64 (n,H,W)."""
    g = torch.Generator().manual_seed(seed)
    gh, gw = H // cell, W // cell
    out = torch.empty((n, H, W), dtype=torch.int64)
    for b in range(n):
        cls = torch.randperm(num_real, generator=g)[:k]
        grid = cls[torch.randint(0, k, (gh, gw), generator=g)]
        ign = torch.rand(gh, gw, generator=g) < ignore_frac
        grid = torch.where(ign, torch.full_like(grid, num_real), grid)
        full = grid.repeat_interleave(cell, 0).repeat_interleave(cell, 1)
        out[b, :full.shape[0], :full.shape[1]] = full
        if full.shape[0] < H or full.shape[1] < W:  # ragged border -> ignore
            out[b, full.shape[0]:, :] = num_real
            out[b, :, full.shape[1]:] = num_real
    return out


def synth_features(n, C, H, W, strides, seed):
    """fp32 NCHW feature maps, one per stride, drawn in scale order from one generator."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(n, C, H // s, W // s, generator=g) for s in strides]


# name -> dict(config for the loss, input recipe).  Mirrors BASELINE.json "configs".
CONFIGS = {
    # cfg-1: DCV2 single-scale (the reference's CPU-runnable case)
    "cfg1": dict(
        loss=dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, max_views_per_class=100,
                  min_views_per_class=5, max_features_total=10000),
        single_scale=True, n=2, C=256, H=512, W=1024, strides=[4], labels="uniform", label_seed=0),
    # cfg-2: HRNet-W48 Cityscapes ms+cs, the headline configuration
    "cfg2": dict(
        loss=dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, scales=4,
                  weights=[1.0, 0.7, 0.4, 0.1], cross_scale_contrast=True, detach_deepest=False,
                  w_high_low=1.0, w_high_mid=1.0, min_views_per_class=5,
                  max_views_per_class=2500, max_features_total=10000),
        single_scale=False, n=12, C=256, H=512, W=1024, strides=[4, 8, 16, 32],
        labels=dict(num_real=19, k=19, cell=32, ignore_frac=0.05, seed=0), feat_seed=1),
    # cfg-3: UPerNet Swin-T ADE20K ms+cs
    "cfg3": dict(
        loss=dict(dataset="ADE20K", experiment=1, temperature=0.1, scales=4,
                  weights=[1.0, 0.7, 0.4, 0.1], cross_scale_contrast=True, detach_deepest=False,
                  w_high_low=1.0, w_high_mid=1.0, min_views_per_class=5,
                  max_views_per_class=2500, max_features_total=10000),
        single_scale=False, n=16, C=256, H=512, W=512, strides=[4, 8, 16, 32],
        labels=dict(num_real=150, k=10, cell=32, ignore_frac=0.05, seed=0), feat_seed=1),
    # cfg-4: DeepLabv3 R101 CaDIS single-scale, large anchor budget
    "cfg4": dict(
        loss=dict(dataset="CADIS", experiment=2, temperature=0.1, min_views_per_class=5,
                  max_views_per_class=1000, max_features_total=10000),
        single_scale=True, n=8, C=256, H=544, W=960, strides=[8],
        labels=dict(num_real=17, k=8, cell=8, ignore_frac=0.05, seed=0), feat_seed=1),
    "cfg4_large": dict(
        loss=dict(dataset="CADIS", experiment=2, temperature=0.1, min_views_per_class=5,
                  max_views_per_class=1000, max_features_total=32768),
        single_scale=True, n=8, C=256, H=544, W=960, strides=[8],
        labels=dict(num_real=17, k=8, cell=8, ignore_frac=0.05, seed=0), feat_seed=1),
    # cfg-5: pooled cross-batch anchors (sharded over 2/4/8 GPUs)
    "cfg5": dict(
        loss=dict(dataset="CITYSCAPES", experiment=1, temperature=0.1, scales=4,
                  weights=[1.0, 0.7, 0.4, 0.1], cross_scale_contrast=True, detach_deepest=False,
                  w_high_low=1.0, w_high_mid=1.0, min_views_per_class=5,
                  max_views_per_class=2500, max_features_total=65536),
        single_scale=False, n=64, C=256, H=512, W=1024, strides=[4, 8, 16, 32],
        labels=dict(num_real=19, k=19, cell=32, ignore_frac=0.05, seed=0), feat_seed=1),
}


def make_labels(cfg):
    spec = cfg["labels"]
    if spec == "uniform":
        raise ValueError("cfg1 draws labels and features from the default generator; use make_cfg1_inputs")
    return synth_labels(cfg["n"], cfg["H"], cfg["W"], spec["num_real"], spec["k"], spec["cell"],
                        spec["ignore_frac"], spec["seed"])


def make_cfg1_inputs():
    """cfg-1 as the survey generated it: after ``torch.manual_seed(0)`` draw features then labels
    from the default generator; the loss is then called *without* reseeding."""
    cfg = CONFIGS["cfg1"]
    torch.manual_seed(0)
    feats = torch.randn(cfg["n"], cfg["C"], cfg["H"] // 4, cfg["W"] // 4)
    labels = torch.randint(0, 20, (cfg["n"], cfg["H"], cfg["W"]))
    return labels, feats


def make_inputs(name, with_features=True):
    """Return ``(labels int64 (n,H,W), [features fp32 NCHW per scale])`` on the CPU."""
    cfg = CONFIGS[name]
    if name == "cfg1":
        labels, feats = make_cfg1_inputs()
        return labels, [feats]
    labels = make_labels(cfg)
    feats = synth_features(cfg["n"], cfg["C"], cfg["H"], cfg["W"], cfg["strides"], cfg["feat_seed"]) \
        if with_features else None
    return labels, feats
