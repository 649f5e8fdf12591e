"""Import shim: the package directory is named ``gpu-pruner_b200`` (not a Python identifier),
so ``import gpu_pruner_b200`` resolves here and loads that directory as the package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu-pruner_b200")
_spec = importlib.util.spec_from_file_location(
    "gpu_pruner_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gpu_pruner_b200"] = _mod
_spec.loader.exec_module(_mod)
