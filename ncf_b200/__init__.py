"""Importable alias of `neural-collaborative-filtering-demo_b200/` (a hyphen is not a valid module name)."""
import os as _os

_pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "neural-collaborative-filtering-demo_b200")
__path__.insert(0, _pkg)
with open(_os.path.join(_pkg, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg, "__init__.py"), "exec"))
