"""Import shim: loads the package directory `volumetric-path-tracer_b200/` (hyphenated, hence not a valid
Python identifier) under the module name `vpt_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "volumetric-path-tracer_b200")
_spec = importlib.util.spec_from_file_location("vpt_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vpt_b200"] = _mod
_spec.loader.exec_module(_mod)
