"""Importable alias of the product package, whose directory name (`kaolin-wisp_b200`) is not a Python identifier."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kaolin-wisp_b200"))
_init = _os.path.join(__path__[0], "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init
