"""Importable alias of ``hyres-residual-enhanced-hybrid-image-compression_b200/``.

The package directory carries the repository's full (hyphenated) name; Python
cannot import that spelling, so this shim points ``hyres_b200.__path__`` at it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "hyres-residual-enhanced-hybrid-image-compression_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _fh
