"""extendablefem.jl_b200 -- B200-native assembly engine behind ExtendableFEM.jl's operator API.

The directory name contains a dot, so it is loaded through ``__graft_entry__.load_package()``
(importlib) under the module name ``extfem_b200``.  Contents:

  csrc/   CUDA kernels + the C-ABI (libextfem_cuda.so, declared in include/extfem_cuda.h)
  host/   Python side above the C-ABI (Julia is not in this image): lib.py = ctypes twin of the ccall stubs; grids.py /
          fespace.py = stand-ins for the ExtendableGrids / ExtendableFEMBase containers the path reads; problem.py = mirror
          of ProblemDescription / assign_operator! / BilinearOperator / LinearOperator / NonlinearOperator /
          HomogeneousBoundaryData / InterpolateBoundaryData / ItemIntegrator / assemble_system! / solve / evaluate for
          registry kernels; dist.py = cell partitioning and interface plans of the sharded system
  julia/  the ccall glue a maintainer adds on the Julia side (untested here: no Julia in the image)
"""
from .host import *          # noqa: F401,F403
from .host import __all__    # noqa: F401
