"""Class-count facts the loss needs from the reference's dataset tables.

The reference derives ``num_all_classes`` / ``ignore_class`` from
``utils.DATASETS_INFO[dataset].CLASS_INFO[experiment][1]`` (an id->name dict):
``losses/DenseContrastiveLossV2.py:16-18``.  Only two facts per (dataset, experiment)
matter to the loss: the number of entries and whether key 255 exists.  They are
tabulated here (values computed from the reference tables) so the
loss has no dependency on the reference's ``utils`` package; when that package *is*
importable (the loss is running inside the reference framework) it takes precedence.
"""
import sys

# (dataset, experiment) -> (num_all_classes, has_255_key)
_TABLE = {
    ("CADIS", 0): (36, False), ("CADIS", 1): (8, False), ("CADIS", 2): (18, True), ("CADIS", 3): (26, True),
    ("CITYSCAPES", 0): (37, False), ("CITYSCAPES", 1): (20, True),
    ("PASCALC", 0): (60, False), ("PASCALC", 1): (60, True),
    ("ADE20K", 0): (151, False), ("ADE20K", 1): (151, True),
}


def class_facts(dataset, experiment):
    """Return ``(num_all_classes, num_real_classes, ignore_class)`` as V2.py:16-18 computes them."""
    utils_mod = sys.modules.get("utils")
    info = getattr(utils_mod, "DATASETS_INFO", None) if utils_mod is not None else None
    if info is not None:
        try:
            id2name = info[dataset].CLASS_INFO[experiment][1]
            n_all, has255 = len(id2name), 255 in id2name
        except Exception:  # fall through to the table
            info = None
    if info is None:
        try:
            n_all, has255 = _TABLE[(dataset, int(experiment))]
        except KeyError:
            raise KeyError(f"unknown dataset/experiment {dataset!r}/{experiment!r}; "
                           f"known: {sorted(_TABLE)}") from None
    n_real = n_all - 1 if has255 else n_all
    ignore = n_all - 1 if has255 else -1
    return n_all, n_real, ignore
