"""Shim of the torch_geometric symbols the reference imports (see ../README.md)."""
from . import typing, utils, nn  # noqa: F401
__version__ = "2.0.3-shim"
