"""Import shim: exposes the hyphenated product directory `vae-cyclegan-implementation_b200/` as the
importable package `vcg_b200` (``import vcg_b200`` from the repo root)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vae-cyclegan-implementation_b200")
_spec = importlib.util.spec_from_file_location("vcg_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vcg_b200"] = _mod
_spec.loader.exec_module(_mod)
