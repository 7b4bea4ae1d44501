"""Import shim: ``import pn2_b200`` loads the package directory
``open3d-pointnet2-semantic3d_b200/`` (whose name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "open3d-pointnet2-semantic3d_b200")
_spec = importlib.util.spec_from_file_location(
    "pn2_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pn2_b200"] = _mod
_spec.loader.exec_module(_mod)
