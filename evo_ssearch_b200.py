"""Import shim: the package directory is ``evo-ssearch_b200/`` (named after the reference repo), which
is not a valid Python identifier.  This module makes it importable as ``evo_ssearch_b200``: it turns
itself into a package whose search path is that directory and then runs the package's ``__init__``.
"""
import os as _os

__package__ = __name__
__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "evo-ssearch_b200")]
if globals().get("__spec__") is not None:
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
