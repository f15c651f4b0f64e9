"""mscs_b200 -- B200-native multi-scale / cross-scale dense supervised contrastive loss.

Drop-in for the reference's ``DenseContrastiveLossV2`` / ``DenseContrastiveLossV2_ms``
(losses/DenseContrastiveLossV2.py, losses/DenseContrastiveLossV2_ms.py), implemented as
hand-written sm_100a CUDA (libmscs.so, C ABI in include/mscs.h) behind one autograd.Function.
"""
from . import synth  # noqa: F401
from ._lib import EXPORTS, LIB_PATH, load  # noqa: F401
from .losses import DenseContrastiveLossV2, DenseContrastiveLossV2_ms, install_into_reference  # noqa: F401
from ._ops import CompactLabels, TorchDistComm, ThreadComm, shard_rows  # noqa: F401
from .projector import ProjectorTailContrastive_ms  # noqa: F401
from .coloss import CrossEntropyLabelPass, FusedCoLosses, label_pass  # noqa: F401

__all__ = ["DenseContrastiveLossV2", "DenseContrastiveLossV2_ms", "install_into_reference", "synth", "load"]
