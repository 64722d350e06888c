"""Public surface of viditq_b200."""
from . import _lib, ops  # noqa: F401
from ._lib import VqError, LIB_PATH  # noqa: F401
from .build import build as build_library  # noqa: F401

__all__ = ["ops", "VqError", "LIB_PATH", "build_library"]
