"""Import shim: the product package lives in the directory ``ekf-slam_b200/`` (the name the
build contract fixes); a hyphen cannot appear in a Python module name, so ``import
ekf_slam_b200`` resolves here and this package's search path is pointed at that directory."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "ekf-slam_b200")
__path__.insert(0, _real)

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
