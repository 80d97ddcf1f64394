from .grids import *        # noqa: F401,F403
from .fespace import *      # noqa: F401,F403
from .dist import *         # noqa: F401,F403
from . import grids as _g, fespace as _f, dist as _d, lib, problem  # noqa: F401
__all__ = list(_g.__all__) + list(_f.__all__) + list(_d.__all__) + ["lib", "dist", "problem"]
