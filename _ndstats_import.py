"""Import the product package, whose directory name (`ndarray-stats_b200/`) is not a Python identifier.

`load()` registers it in sys.modules as `ndarray_stats_b200` and returns the module.
"""
import importlib.util
import os
import sys

_NAME = "ndarray_stats_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ndarray-stats_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
