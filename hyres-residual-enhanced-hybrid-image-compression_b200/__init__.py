"""B200-native HyRES residual-codec hot path.

Import name: ``hyres_b200`` (a shim package at the repository root extends its
``__path__`` to this directory, whose hyphenated name is not importable).
"""
from . import _lib  # noqa: F401
