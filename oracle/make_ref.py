"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, staged where they can travel.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is a Python tree that lives only in the build
container (``/root/reference``); the GPU box has no such path.  This script copies -- byte for byte, nothing is
edited -- the few files the path needs into the git-ignored directory ``oracle/_ref/`` (it ships to the GPU box with
the repository snapshot exactly like the built ``libmscs.so`` does; it never enters the history):

    losses/*.py                      LossWrapper.py (the boundary, :33,:68-71,:90), DenseContrastiveLossV2.py,
                                     DenseContrastiveLossV2_ms.py and the two co-loss files ``losses/__init__.py`` imports
    utils/defaults.py                DATASETS_INFO (class tables, DenseContrastiveLossV2.py:16-18)
    utils/datasets_info/*.py         pure-Python dataset tables

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` exists; ``oracle/ref_loader.py`` imports the staged
files (or ``/root/reference`` itself) under a stub ``utils`` package (the real ``utils/__init__.py`` needs matplotlib
and cv2).  Used by: tests (the GPU test that drives the unmodified ``LossWrapper.forward`` + backward with this
repository's classes installed), ``bench.py --impl reference`` (the reference's own CPU implementation, timed).
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MSCS_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["losses/__init__.py", "losses/LossWrapper.py", "losses/DenseContrastiveLossV2.py",
         "losses/DenseContrastiveLossV2_ms.py", "losses/LovaszSoftmax.py", "losses/TwoScaleLoss.py",
         "utils/defaults.py", "utils/datasets_info/__init__.py", "utils/datasets_info/CITYSCAPES.py",
         "utils/datasets_info/CADIS.py", "utils/datasets_info/PASCALC.py", "utils/datasets_info/ADE20K.py"]


def stage(verbose=True):
    """Copies FILES from the reference checkout into oracle/_ref/.  Returns False when there is no checkout."""
    if not os.path.isdir(os.path.join(REF, "losses")):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    if verbose:
        print(f"staged {len(FILES)} unmodified reference files from {REF} into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
