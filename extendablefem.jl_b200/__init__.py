"""extendablefem.jl_b200 -- B200-native assembly engine behind ExtendableFEM.jl's operator API.

The directory name contains a dot, so it is loaded through ``__graft_entry__.load_package()``
(importlib) under the module name ``extfem_b200``.  Contents:

  csrc/   CUDA kernels + the C-ABI (libextfem_cuda.so, declared in include/extfem_cuda.h)
  host/   Python mirror of the reference's operator API (ProblemDescription, BilinearOperator,
          LinearOperator, NonlinearOperator, assemble!, solve) on top of the C-ABI via ctypes
  julia/  the ccall glue a maintainer adds on the Julia side (untested here: no Julia in the image)
"""
from .host import *          # noqa: F401,F403
from .host import __all__    # noqa: F401
