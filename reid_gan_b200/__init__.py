"""Import alias: the product package lives in the directory `reid-gan_b200/` (the
name the build contract fixes), which is not a valid Python identifier.  This
shim makes it importable as `reid_gan_b200` by pointing the package search path
at that directory and executing its __init__.py."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "reid-gan_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
