"""Importable alias of the `vidit-q_b200/` package directory (a hyphen is not a legal module name).

`import viditq_b200` resolves submodules (`viditq_b200.ops`, `viditq_b200.qdiff`, ...) from ../vidit-q_b200/.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vidit-q_b200"))

from ._api import *  # noqa: E402,F401,F403
