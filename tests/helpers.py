"""Shared helpers of the test-suite (oracle-side config, golden loading, comparisons)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# num_all_classes per (dataset, experiment) -- SURVEY.md §8a
CLASSES = {("CITYSCAPES", 1): 20, ("ADE20K", 1): 151, ("CADIS", 1): 8, ("CADIS", 2): 18, ("CADIS", 3): 26}


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def oracle_cfg_for(meta):
    from oracle.config import oracle_cfg
    lc = meta["loss_cfg"]
    cfg = oracle_cfg(lc, CLASSES[(lc["dataset"], lc["experiment"])])
    if meta["single_scale"]:
        cfg["cross_scale"] = False
        cfg["weights"] = [1.0]
    return cfg


def small_case_inputs(name):
    """labels (int64) and feature list (fp32) of a fully stored golden case."""
    z = load_npz(name)
    feats = []
    s = 0
    while f"feat{s}" in z:
        feats.append(torch.from_numpy(z[f"feat{s}"]))
        s += 1
    return torch.from_numpy(z["labels"].astype(np.int64)), feats, z


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / np.sqrt((a @ a) * (b @ b)))


def make_module(meta, device=None):
    import mscs_b200
    cls = mscs_b200.DenseContrastiveLossV2 if meta["single_scale"] else mscs_b200.DenseContrastiveLossV2_ms
    return cls(dict(meta["loss_cfg"]))
