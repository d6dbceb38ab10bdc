"""Importable alias for the package whose sources live in ``gnn-gp_b200/``.

The repo layout names the package directory ``gnn-gp_b200`` (a hyphen is not a
legal Python identifier), so this stub extends its ``__path__`` to that
directory: ``import gnngp_b200.compute`` resolves to ``gnn-gp_b200/compute.py``.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gnn-gp_b200")
if not _os.path.isdir(_SRC):  # pragma: no cover
    raise ImportError("gnngp_b200: source directory %s is missing" % _SRC)
__path__.insert(0, _SRC)

from ._version import __version__  # noqa: E402,F401
