"""Importable alias of the product package.

The product lives in ``task-specific-pretraining-multimodal_b200/`` (a directory name that is not a valid Python
identifier); this shim makes it importable as ``mml_b200`` without copying anything.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "task-specific-pretraining-multimodal_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
