"""Import alias for the product package.

The package directory carries the reference repo's name
(``eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200``), which is not a
valid Python identifier.  ``import mscs_b200`` loads that directory as the package
``mscs_b200`` (relative imports inside it work as usual).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200")
_spec = _ilu.spec_from_file_location("mscs_b200", _os.path.join(_PKG_DIR, "__init__.py"),
                                     submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["mscs_b200"] = _mod
_spec.loader.exec_module(_mod)
