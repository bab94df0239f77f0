"""Importable alias of the ``indoor-nerf_b200/`` package directory (a hyphen cannot appear in an import
statement).  All code lives in ``indoor-nerf_b200/``; this file only points the package path there and
runs that directory's ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "indoor-nerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
