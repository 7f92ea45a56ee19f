"""Import shim: the package directory is named after the reference repo (`icl-mixed-precision-gmres_b200/`), which is
not a valid Python identifier; this module loads it under the name `icl_mixed_precision_gmres_b200`.

    import gmres_b200 as g
    ctx = g.Context(0)
"""
import importlib.util
import os
import sys

_NAME = "icl_mixed_precision_gmres_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icl-mixed-precision-gmres_b200")
if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
pkg = sys.modules[_NAME]
from icl_mixed_precision_gmres_b200 import *  # noqa: E402,F401,F403
PACKAGE_DIR = _DIR
