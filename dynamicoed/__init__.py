"""Import alias: makes the on-disk package directory ``dynamicoed.jl_b200/`` importable as
``dynamicoed.jl_b200`` (a directory name containing a dot cannot be imported directly)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "dynamicoed.jl_b200")
if "dynamicoed.jl_b200" not in _sys.modules:
    _spec = _ilu.spec_from_file_location("dynamicoed.jl_b200",
                                         _os.path.join(_dir, "__init__.py"),
                                         submodule_search_locations=[_dir])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["dynamicoed.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = _sys.modules["dynamicoed.jl_b200"]
