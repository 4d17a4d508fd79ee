"""Import shim: the product package directory is named after the reference repository and is not a
valid Python identifier, so it is registered here under the importable name `nerfq_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200")
_spec = importlib.util.spec_from_file_location("nerfq_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nerfq_b200"] = _mod
_spec.loader.exec_module(_mod)
