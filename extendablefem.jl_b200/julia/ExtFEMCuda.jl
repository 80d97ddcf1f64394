# ExtFEMCuda.jl -- ccall glue between ExtendableFEM.jl and libextfem_cuda.so (include/extfem_cuda.h).
#
# UNTESTED in this repository: Julia is not part of the build image.  It is the file a maintainer adds as
# ext/ExtendableFEMCudaExt.jl; INTEGRATION.md explains every call and cites the reference lines it replaces
# (O.assembler closures built in src/common_operators/*_operator.jl: build_assembler!).
module ExtFEMCuda
using ExtendableFEM, ExtendableFEMBase, ExtendableGrids, SparseArrays
const lib = "libextfem_cuda"            # extendablefem.jl_b200/csrc/libextfem_cuda.so on LD_LIBRARY_PATH

# mirror of `extfem_opdesc` (include/extfem_cuda.h:92-114); NTuple{4,Int32} == int32_t[EXTFEM_MAXARGS]
struct OpDesc
    ntest::Int32;   test_block::NTuple{4,Int32};   test_op::NTuple{4,Int32}
    nansatz::Int32; ansatz_block::NTuple{4,Int32}; ansatz_op::NTuple{4,Int32}
    nargs::Int32;   args_block::NTuple{4,Int32};   args_op::NTuple{4,Int32}
    kernel_id::Int32; nparams::Int32; params::Ptr{Float64}
    factor::Float64; time::Float64; symgrad_offdiag::Float64
    quadorder::Int32; bonus_quadorder::Int32
    nregions::Int32; regions::Ptr{Int32}
    transposed_copy::Int32; lump::Int32; coupling::Ptr{UInt8}
    nq_custom::Int32; qweights::Ptr{Float64}; qpoints::Ptr{Float64}; tabulated::Ptr{Float64}
end

check(ctx, rc) = rc == 0 || error(unsafe_string(ccall((:extfem_last_error, lib), Cstring, (Ptr{Cvoid},), ctx)))

mutable struct Context; ptr::Ptr{Cvoid}; end
function Context(device = 0)
    r = Ref{Ptr{Cvoid}}()
    check(C_NULL, ccall((:extfem_ctx_create, lib), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, r))
    finalizer(c -> ccall((:extfem_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.ptr), Context(r[]))
end

# registry: Julia function object -> kernel id; an unregistered closure is an error (north_star)
const REGISTRY = IdDict{Any,String}(ExtendableFEMBase.standard_kernel => "standard",
                                    ExtendableFEM.constant_one_kernel => "constant_one")
register_kernel!(f, name) = (REGISTRY[f] = name)
function kernel_id(f)
    haskey(REGISTRY, f) || error("ExtFEMCuda: kernel $(f) is not registered with the GPU engine (no CPU fallback)")
    id = ccall((:extfem_kernel_id, lib), Cint, (Cstring,), REGISTRY[f]); id > 0 || error("unknown kernel"); id
end

# xgrid[Coordinates], xgrid[CellNodes], xgrid[CellRegions], xgrid[CellVolumes]  (bilinear_operator.jl:693-695)
function mesh_set(ctx, xgrid::ExtendableGrid{Tv,Ti}) where {Tv,Ti}
    X, CN = xgrid[Coordinates], xgrid[CellNodes]; R = Int32.(xgrid[CellRegions]); V = xgrid[CellVolumes]
    h = Ref{Cint}()
    GC.@preserve X CN R V check(ctx.ptr, ccall((:extfem_mesh_set, lib), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Cint}),
        ctx.ptr, size(X, 1), size(CN, 2), size(X, 2), X, CN, sizeof(Ti), R, V, h))
    h[]
end

# FES[CellDofs] via get_dofmap (helper_functions.jl:561-567); VariableTargetAdjacency -> dense [nd, ncells]
function space_set(ctx, mesh, FES::FESpace{Tv,Ti,FEType}) where {Tv,Ti,FEType}
    CD = Matrix(FES[CellDofs]); fe = FEType <: H1P1 ? 1 : FEType <: H1P2 ? 2 :
         (FEType <: H1Pk && get_polynomialorder(FEType, Tetrahedron3D) <= 2) ? get_polynomialorder(FEType, Tetrahedron3D) :
         error("ExtFEMCuda: $FEType is not supported (EXTFEM_ERR_UNSUPPORTED_ELEMENT)")
    h = Ref{Cint}()
    GC.@preserve CD check(ctx.ptr, ccall((:extfem_space_set, lib), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Cint}),
        ctx.ptr, mesh, fe, get_ncomponents(FEType), CD, sizeof(Ti), size(CD, 1), FES.ndofs, h))
    h[]
end

# pattern once per FEMatrix; colptr/rowval come back Int64 1-based, rows sorted: drop straight into SparseMatrixCSC
function pattern!(ctx, A::FEMatrix, spaces)
    h = Ref{Cint}(); s = Cint.(spaces)
    check(ctx.ptr, ccall((:extfem_pattern_build, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Ptr{UInt8}, Ptr{Cint}),
        ctx.ptr, length(s), s, length(s), s, C_NULL, h))
    nr, nc, nnz = Ref{Int64}(), Ref{Int64}(), Ref{Int64}()
    ccall((:extfem_pattern_dims, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), ctx.ptr, h[], nr, nc, nnz)
    colptr, rowval = Vector{Int64}(undef, nc[] + 1), Vector{Int64}(undef, nnz[])
    ccall((:extfem_pattern_get, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}), ctx.ptr, h[], colptr, rowval)
    A.entries.cscmatrix = SparseMatrixCSC(nr[], nc[], colptr, rowval, zeros(nnz[]))   # flush!ed by construction
    h[]
end

# the replacement closure: same call shape as O.assembler(A.entries, b.entries)
function gpu_assembler(ctx, pattern, O::BilinearOperator, desc::OpDesc)
    return function (A, b; accumulate = true)          # assemble_system! zeroed nzval already (solvers.jl:134)
        nz = A.cscmatrix.nzval
        GC.@preserve nz check(ctx.ptr, ccall((:extfem_assemble_bilinear, lib), Cint,
            (Ptr{Cvoid}, Cint, Ref{OpDesc}, Ptr{Float64}, Cint, Ptr{Float64}), ctx.ptr, pattern, desc, C_NULL, accumulate, nz))
    end
end
# LinearOperator / NonlinearOperator: identical shape with extfem_assemble_linear / extfem_assemble_nonlinear
# (sol = sol.entries, b_out = b.entries); HomogeneousData.apply_penalties! -> extfem_apply_penalties;
# compute_nonlinear_residual! (solvers.jl:38-43) -> extfem_residual.
end
