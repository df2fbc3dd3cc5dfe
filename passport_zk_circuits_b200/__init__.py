"""Import shim: the package directory is `passport-zk-circuits_b200/` (the name the
project layout prescribes), which is not a valid Python identifier.  This
module makes it importable as `passport_zk_circuits_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "passport-zk-circuits_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
