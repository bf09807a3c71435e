"""Import shim: ``import mm_interleaved_b200`` -> the package directory ``mm-interleaved_b200/``.

The package directory carries the project's hyphenated name, which is not a valid Python
identifier; this one-file shim registers it under the importable name.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mm-interleaved_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
