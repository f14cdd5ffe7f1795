"""Importable alias of the package directory `concrete-numpy_b200/` (a hyphen cannot be imported).

`import concrete_numpy_b200` executes `concrete-numpy_b200/__init__.py` with this module's
`__path__` pointing at that directory, so `concrete_numpy_b200.<submodule>` resolves there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "concrete-numpy_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
