"""Import shim: the product package lives in the directory `juliagrid.jl_b200/` (a name Python's import
statement cannot spell), and is importable as `jgb200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "juliagrid.jl_b200")
_spec = importlib.util.spec_from_file_location("jgb200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["jgb200"] = _mod
_spec.loader.exec_module(_mod)
