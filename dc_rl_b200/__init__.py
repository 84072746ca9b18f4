"""Importable alias for the package directory `dc-rl_b200/` (a hyphen cannot appear in a Python
module name). All code lives in `dc-rl_b200/`; this file only points the import system at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dc-rl_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
