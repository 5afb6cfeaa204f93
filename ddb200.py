"""Import shim: the product package lives in ``datadrivendiffeq.jl_b200/`` (the name the project
layout fixes); a dot is not legal in a Python package name, so ``import ddb200`` loads that
directory as the package ``ddb200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "datadrivendiffeq.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "ddb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ddb200"] = _mod
_spec.loader.exec_module(_mod)
