"""Import shim: the package directory is ``biped-pympc_b200/`` (hyphen, not importable by
name); ``import biped_pympc_b200`` loads it under this module name."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "biped-pympc_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
