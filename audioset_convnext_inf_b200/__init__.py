"""Importable alias for the package directory `audioset-convnext-inf_b200/` (a hyphen cannot appear
in a Python module name).  `import audioset_convnext_inf_b200` executes that directory's package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "audioset-convnext-inf_b200")
__path__[:] = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
del _fh, _os, _real
