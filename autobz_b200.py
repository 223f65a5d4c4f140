"""Import shim: the package directory is named `autobzcore.jl_b200` (not a valid Python identifier),
so `import autobz_b200` loads it under this name."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "autobzcore.jl_b200")
_spec = importlib.util.spec_from_file_location("autobz_b200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["autobz_b200"] = _mod
_spec.loader.exec_module(_mod)
