"""Import shim: exposes the package directory `wave-fenics_b200/` as `wave_fenics_b200`."""
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "wave-fenics_b200")
__path__ = [_pkg_dir]
__package__ = __name__
__file__ = _os.path.join(_pkg_dir, "__init__.py")
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"), globals())
