"""Import shim: makes the on-disk package directory `lfpsqp.jl_b200/` importable as `lfpsqp.jl_b200`."""
import importlib.util
import os
import sys

_pkgdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lfpsqp.jl_b200")
_name = "lfpsqp.jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_pkgdir, "__init__.py"),
                                                   submodule_search_locations=[_pkgdir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = sys.modules[_name]
