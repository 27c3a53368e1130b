"""Import alias: ``import ciri_long_b200`` -> the package in ./ciri-long_b200/ (whose directory name
contains a hyphen and therefore cannot be imported by name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ciri-long_b200")
_spec = importlib.util.spec_from_file_location(
    "ciri_long_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ciri_long_b200"] = _mod
_spec.loader.exec_module(_mod)
