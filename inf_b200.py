"""Import shim: ``import inf_b200`` loads the package that lives in ``intrinsic-neural-fields_b200/``
(the directory name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "intrinsic-neural-fields_b200")
_spec = importlib.util.spec_from_file_location(
    "inf_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["inf_b200"] = _mod
_spec.loader.exec_module(_mod)
