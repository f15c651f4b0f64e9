"""Import of the UNMODIFIED reference loss modules -- TEST INFRASTRUCTURE (see oracle/__init__.py).

``load()`` imports the reference's ``losses`` package (LossWrapper.py, DenseContrastiveLossV2.py,
DenseContrastiveLossV2_ms.py, ...) from ``/root/reference`` (build container) or from the staged copy
``oracle/_ref/`` (GPU box; recipe: oracle/make_ref.py).  Two shims, no file edits:

  * a stub ``utils`` package serving the real ``utils.defaults.DATASETS_INFO`` (the real ``utils/__init__.py`` pulls
    matplotlib / cv2, absent here) plus the five helpers the loss files import (rank 0, no-op logging);
  * ``cpu=True`` only: ``torch.Tensor.cuda`` -> identity for the hard-coded ``.cuda()`` calls
    (DenseContrastiveLossV2.py:113,114,121,168) so that the reference runs on the host cores.  On a GPU box the
    reference runs as it is (``cpu=False``).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get("MSCS_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref")]


def find_root():
    for root in CANDIDATES:
        if root and os.path.isfile(os.path.join(root, "losses", "LossWrapper.py")) and \
                os.path.isfile(os.path.join(root, "utils", "defaults.py")):
            return root
    return None


_loaded = {}


def load(cpu):
    """Returns a namespace: root, losses (the reference package), LossWrapper, DenseContrastiveLossV2,
    DenseContrastiveLossV2_ms (the reference CLASSES, captured before anything is installed over them), DATASETS_INFO.
    Raises FileNotFoundError when neither location holds the files."""
    import torch
    if "ns" not in _loaded:
        root = find_root()
        if root is None:
            raise FileNotFoundError("reference files not found: neither /root/reference nor oracle/_ref/ "
                                    "(run `python oracle/make_ref.py` in the build container)")
        for name in ("utils", "losses"):
            if name in sys.modules:
                raise RuntimeError(f"a module named {name!r} is already imported; the reference needs that name")
        sys.path.insert(0, root)
        utils = types.ModuleType("utils")
        utils.__path__ = [os.path.join(root, "utils")]
        sys.modules["utils"] = utils
        defaults = importlib.import_module("utils.defaults")       # the real dataset tables (pure Python)
        utils.DATASETS_INFO = defaults.DATASETS_INFO
        utils.get_rank = lambda: 0
        utils.printlog = lambda *a, **k: None
        utils.is_distributed = lambda: False
        utils.concat_all_gather = None
        utils.to_numpy = lambda t: t.detach().cpu().numpy()
        utils.Logger = types.SimpleNamespace(info=lambda *a, **k: None)
        losses = importlib.import_module("losses")                  # the reference package, unmodified
        _loaded["ns"] = types.SimpleNamespace(
            root=root, losses=losses, wrapper_module=sys.modules["losses.LossWrapper"],
            LossWrapper=losses.LossWrapper, DenseContrastiveLossV2=losses.DenseContrastiveLossV2,
            DenseContrastiveLossV2_ms=losses.DenseContrastiveLossV2_ms, DATASETS_INFO=defaults.DATASETS_INFO)
    if cpu and not _loaded.get("cpu"):
        torch.Tensor.cuda = lambda self, *a, **k: self
        _loaded["cpu"] = True
    return _loaded["ns"]
