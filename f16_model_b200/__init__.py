"""Importable alias of the package directory ``f16-model_b200/`` (a hyphen is not a
valid Python identifier).  ``import f16_model_b200`` loads ``f16-model_b200/__init__.py``
under this name; sub-modules resolve inside the hyphenated directory."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "f16-model_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
