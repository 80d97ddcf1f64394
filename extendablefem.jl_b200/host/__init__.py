from .grids import *        # noqa: F401,F403
from .fespace import *      # noqa: F401,F403
from . import grids as _g, fespace as _f, lib  # noqa: F401
__all__ = list(_g.__all__) + list(_f.__all__) + ["lib"]
