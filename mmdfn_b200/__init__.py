"""Importable alias of the package directory ``mm-dfn_b200/`` (a hyphen is not a legal
Python identifier).  ``import mmdfn_b200`` loads ``mm-dfn_b200/__init__.py`` under this
name; all sub-modules resolve inside that directory."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mm-dfn_b200")
_spec = _ilu.spec_from_file_location("mmdfn_b200", _os.path.join(_real, "__init__.py"),
                                     submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["mmdfn_b200"] = _mod
_spec.loader.exec_module(_mod)
