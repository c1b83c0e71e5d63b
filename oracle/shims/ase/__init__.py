"""Minimal ASE stand-in (oracle/shims/README.md). Not ASE."""
from . import atoms  # noqa: F401
from .atoms import Atoms  # noqa: F401

__version__ = "0.0-shim"
