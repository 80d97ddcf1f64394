# ExtFEMCuda.jl -- ccall glue between ExtendableFEM.jl and libextfem_cuda.so (include/extfem_cuda.h).
#
# NOT EXECUTED in this repository: Julia is not part of the build image (probed: `julia` absent).  The same calls, in the same
# order and with the same argument meaning, are exercised through the ctypes twin (host/lib.py + host/problem.py) by the
# GPU tests and through a plain C caller fed with Julia-layout arrays (tests/abi_driver.c).  A maintainer adds this file as
# ext/ExtendableFEMCudaExt.jl; INTEGRATION.md walks through it.
#
# Seam: the closure `O.assembler` that build_assembler! creates (bilinear_operator.jl:955-1003, linear_operator.jl:642-695,
# nonlinear_operator.jl:440-477) and assemble! invokes with raw arrays (:1031 / :713 / :491).  `gpu!(O, gs)` installs a
# closure with the same call shape that ADDS the operator's contribution into the arrays it is handed -- exactly what the
# reference closure does -- so `solve`, `ProblemDescription`, `assign_operator!` and `assemble_system!` stay untouched.
module ExtFEMCuda
using ExtendableFEM, ExtendableFEMBase, ExtendableGrids, SparseArrays, LinearAlgebra
const lib = "libextfem_cuda"            # extendablefem.jl_b200/csrc/libextfem_cuda.so on LD_LIBRARY_PATH

# ---- mirror of `extfem_opdesc` (include/extfem_cuda.h); NTuple{4,Int32} == int32_t[EXTFEM_MAXARGS] -----------------------
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
    entities::Int32
end

const ON_CELLS_ID, ON_BFACES_ID = Int32(0), Int32(1)
check(ctx, rc) = rc == 0 || error(unsafe_string(ccall((:extfem_last_error, lib), Cstring, (Ptr{Cvoid},), ctx)))

mutable struct Context; ptr::Ptr{Cvoid}; end
function Context(device = 0)
    r = Ref{Ptr{Cvoid}}()
    check(C_NULL, ccall((:extfem_ctx_create, lib), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, r))
    finalizer(c -> ccall((:extfem_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.ptr), Context(r[]))
end

# ---- registry: Julia function object -> registry name; an unregistered closure is an error (north_star: no fallback) -----
const REGISTRY = IdDict{Any,String}(ExtendableFEMBase.standard_kernel => "standard",
                                    ExtendableFEMBase.constant_one_kernel => "constant_one",
                                    ExtendableFEM.l2norm_kernel => "l2norm")
register_kernel!(f, name::String) = (REGISTRY[f] = name)      # e.g. register_kernel!(Example250.kernel_nonlinear!, "nse2d")
function kernel_id(f)
    haskey(REGISTRY, f) || error("ExtFEMCuda: kernel $(f) is not registered with the GPU engine (no CPU fallback)")
    id = ccall((:extfem_kernel_id, lib), Cint, (Cstring,), REGISTRY[f])
    id > 0 || error("ExtFEMCuda: registry name $(REGISTRY[f]) is unknown to libextfem_cuda")
    return id
end

opcode(::Type{<:Identity}) = Int32(0)
opcode(::Type{<:Gradient}) = Int32(1)
opcode(::Type{<:Divergence}) = Int32(2)
opcode(::Type{<:SymmetricGradient}) = Int32(3)
opcode(T) = error("ExtFEMCuda: function operator $T is not supported (EXTFEM_ERR_UNSUPPORTED_ELEMENT)")
offdiag(::Type{SymmetricGradient{v}}) where {v} = Float64(v)
offdiag(_) = 1.0
pad4(v) = NTuple{4,Int32}(ntuple(i -> i <= length(v) ? Int32(v[i]) : Int32(0), 4))

# ---- one device-resident system per (grid, FESpaces): mesh, spaces, pattern handles ------------------------------------
mutable struct GPUSystem
    ctx::Context
    mesh::Cint
    spaces::Vector{Cint}
    pattern::Cint
    nnz::Int64
    nrows::Int64
    stage_nz::Vector{Float64}     # staging of one operator's contribution (added into the caller's arrays)
    stage_b::Vector{Float64}
end

# xgrid[Coordinates], [CellNodes], [CellRegions], [CellVolumes] (bilinear_operator.jl:693-695) + the boundary faces (:707-714)
function mesh_set(ctx, xgrid::ExtendableGrid{Tv,Ti}) where {Tv,Ti}
    X, CN = xgrid[Coordinates], Matrix(xgrid[CellNodes]); R = Int32.(xgrid[CellRegions]); V = xgrid[CellVolumes]
    h = Ref{Cint}()
    GC.@preserve X CN R V check(ctx.ptr, ccall((:extfem_mesh_set, lib), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Cint}),
        ctx.ptr, size(X, 1), size(CN, 2), size(X, 2), X, CN, sizeof(Ti), R, V, h))
    BN = Matrix(xgrid[BFaceNodes]); BR = Int32.(xgrid[BFaceRegions]); BV = xgrid[BFaceVolumes]
    GC.@preserve BN BR BV check(ctx.ptr, ccall((:extfem_mesh_set_bfaces, lib), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}), ctx.ptr, h[], size(BN, 2), BN, sizeof(Ti), BR, BV))
    return h[]
end

# monomial exponents in the order extfem_space_set_tables documents (for k: for j: for i, i fastest)
function monomials(order, dim)
    ex = NTuple{3,Int}[]
    for k in 0:(dim >= 3 ? order : 0), j in 0:(dim >= 2 ? order - k : 0), i in 0:(order - k - j)
        push!(ex, (i, j, k))
    end
    return ex
end

# polynomial coefficients of a scalar reference basis from its values at a unisolvent point set (Vandermonde solve);
# `refbasis!(vals, xref)` is the element's get_basis(ON_CELLS, FEType, EG) closure from ExtendableFEMBase
function basis_coefficients(refbasis!, nscalar, ncomp, order, dim)
    ex = monomials(order, dim)
    pts = [collect(Float64, e[1:dim]) ./ max(order, 1) for e in ex]          # the principal lattice of the simplex
    V = [prod(p[d]^e[d] for d in 1:dim) for p in pts, e in ex]               # [point, monomial]
    B = zeros(length(pts), nscalar); tmp = zeros(nscalar * ncomp, ncomp)
    for (i, p) in enumerate(pts)
        refbasis!(tmp, p); B[i, :] .= tmp[1:nscalar, 1]
    end
    return permutedims(V \ B)                                                 # [nscalar, nmono] row-major after copy
end

# FES[CellDofs] via get_dofmap (helper_functions.jl:561-567); FES[BFaceDofs] for ON_BFACES operators
function space_set(ctx, mesh, FES::FESpace{Tv,Ti,FEType}) where {Tv,Ti,FEType}
    CD = Matrix(FES[CellDofs]); ncomp = get_ncomponents(FEType)
    EG = FES.xgrid[UniqueCellGeometries][1]; order = get_polynomialorder(FEType, EG); dim = dim_element(EG)
    builtin = (FEType <: H1P1) || (FEType <: H1P2) || (FEType <: H1Pk && order <= 2)
    FEType <: Union{H1P1,H1P2,H1Pk} || error("ExtFEMCuda: $FEType is not supported (EXTFEM_ERR_UNSUPPORTED_ELEMENT)")
    fe = builtin ? Cint(order) : Cint(100)                                    # EXTFEM_FE_H1P1 / H1P2 / TABULATED
    if !builtin
        # one reference basis per space: the engine expects the orientation of edge dofs in the dof map, so the per-cell
        # basis permutation ExtendableFEMBase applies (get_basissubset / coefficient handlers) is folded into CD here
        CD = orient_celldofs(FES, CD)
    end
    h = Ref{Cint}()
    GC.@preserve CD check(ctx.ptr, ccall((:extfem_space_set, lib), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Cint}),
        ctx.ptr, mesh, fe, ncomp, CD, sizeof(Ti), size(CD, 1), FES.ndofs, h))
    if !builtin
        ns = size(CD, 1) ÷ ncomp
        C = basis_coefficients(ExtendableFEMBase.get_basis(ON_CELLS, FEType, EG), ns, ncomp, order, dim)
        FG = facetype_of_cellface(EG, 1)
        nsb = size(FES[BFaceDofs], 1) ÷ ncomp
        CB = basis_coefficients(ExtendableFEMBase.get_basis(ON_BFACES, FEType, FG), nsb, ncomp, order, dim - 1)
        Cc, CBc = collect(C'), collect(CB')                                   # row-major [nscalar][nmono]
        GC.@preserve Cc CBc check(ctx.ptr, ccall((:extfem_space_set_tables, lib), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, Cint, Ptr{Float64}), ctx.ptr, h[], order, ns, Cc, nsb, CBc))
    end
    BD = Matrix(FES[BFaceDofs])
    GC.@preserve BD check(ctx.ptr, ccall((:extfem_space_set_bfacedofs, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Cint),
        ctx.ptr, h[], BD, sizeof(Ti), size(BD, 1)))
    return h[]
end

# H1Pk order >= 3: two (or more) dofs per edge whose order follows the edge orientation; swap them in cells that see the edge
# against its global direction (the sign information ExtendableFEMBase keeps in CellFaceSigns / CellEdgeSigns)
function orient_celldofs(FES, CD)
    xgrid = FES.xgrid; CD = copy(CD)
    signs = dim_element(xgrid[UniqueCellGeometries][1]) == 2 ? xgrid[CellFaceSigns] : xgrid[CellEdgeSigns]
    nv = size(xgrid[CellNodes], 1); ne = size(signs, 1)
    nde = (size(CD, 1) ÷ get_ncomponents(eltype(FES)) - nv) ÷ ne                # dofs per edge (cell dofs beyond this are interior)
    nde >= 2 || return CD
    ns = size(CD, 1) ÷ get_ncomponents(eltype(FES))
    for cell in axes(CD, 2), e in 1:ne
        signs[e, cell] < 0 || continue
        for c in 0:(get_ncomponents(eltype(FES)) - 1)
            r = (c * ns + nv + (e - 1) * nde + 1):(c * ns + nv + e * nde)
            CD[r, cell] .= reverse(CD[r, cell])
        end
    end
    return CD
end

# pattern once per FEMatrix; colptr/rowval come back Int64 1-based, rows sorted: drop straight into SparseMatrixCSC
function GPUSystem(ctx::Context, A::FEMatrix, FES::Vector{<:FESpace})
    mesh = mesh_set(ctx, FES[1].xgrid)
    spaces = Cint[space_set(ctx, mesh, F) for F in FES]
    h = Ref{Cint}()
    check(ctx.ptr, ccall((:extfem_pattern_build, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Ptr{UInt8}, Ptr{Cint}),
        ctx.ptr, length(spaces), spaces, length(spaces), spaces, C_NULL, h))
    nr, nc, nnz = Ref{Int64}(), Ref{Int64}(), Ref{Int64}()
    check(ctx.ptr, ccall((:extfem_pattern_dims, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), ctx.ptr, h[], nr, nc, nnz))
    colptr, rowval = Vector{Int64}(undef, nc[] + 1), Vector{Int64}(undef, nnz[])
    check(ctx.ptr, ccall((:extfem_pattern_get, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}), ctx.ptr, h[], colptr, rowval))
    A.entries.cscmatrix = SparseMatrixCSC(nr[], nc[], colptr, rowval, zeros(nnz[]))   # flush!ed by construction (structural pattern)
    return GPUSystem(ctx, mesh, spaces, h[], nnz[], nr[], zeros(nnz[]), zeros(nr[]))
end

# ---- kwargs tables -> OpDesc (bilinear_operator.jl:56-73, linear_operator.jl:34-47, nonlinear_operator.jl:31-48) -------
entity_id(AT) = AT <: ON_CELLS ? ON_CELLS_ID : AT <: ON_BFACES ? ON_BFACES_ID :
                error("ExtFEMCuda: entities = $AT is not supported (ON_CELLS, ON_BFACES)")
function opdesc(P::Dict{Symbol,Any}, kernel, blocks_test, ops_test, blocks_ansatz, ops_ansatz, blocks_args, ops_args; time = 0.0, keep)
    P[:parallel] && @warn "ExtFEMCuda: parallel = true is ignored (the device assembly needs no partition colouring)" maxlog = 1
    get(P, :parallel_groups, false) && @warn "ExtFEMCuda: parallel_groups is ignored" maxlog = 1
    params = Float64.(something(P[:params], Float64[])); regions = Int32.(collect(P[:regions])); push!(keep, params, regions)
    ops = vcat(ops_test, ops_ansatz, ops_args); sg = filter(o -> o <: SymmetricGradient, ops)
    return OpDesc(length(ops_test), pad4(blocks_test .- 1), pad4(opcode.(ops_test)),
                  length(ops_ansatz), pad4(blocks_ansatz .- 1), pad4(opcode.(ops_ansatz)),
                  length(ops_args), pad4(blocks_args .- 1), pad4(opcode.(ops_args)),
                  kernel_id(kernel), length(params), pointer(params), Float64(P[:factor]), Float64(time),
                  isempty(sg) ? 1.0 : offdiag(sg[1]),
                  P[:quadorder] == "auto" ? Int32(-1) : Int32(P[:quadorder]), Int32(P[:bonus_quadorder]),
                  length(regions), pointer(regions), Int32(get(P, :transposed_copy, 0)), Int32(get(P, :lump, 0)), C_NULL,
                  0, C_NULL, C_NULL, C_NULL, entity_id(P[:entities]))
end

# ---- replacement closures: same call shapes as the reference's O.assembler; each ADDS its operator into the caller's arrays --
# BilinearOperator: O.assembler(A.entries, b.entries[, sol blocks]; time)   (bilinear_operator.jl:955, :1031)
function gpu!(O::BilinearOperator, gs::GPUSystem, SC; blocks_test, blocks_ansatz = blocks_test, blocks_args = Int[])
    O.assembler = function (A, b, sol = nothing; time = 0.0, kwargs...)
        keep = Any[]
        d = opdesc(O.parameters, O.kernel, blocks_test, O.ops_test, blocks_ansatz, O.ops_ansatz, blocks_args, O.ops_args; time, keep)
        solv = sol === nothing ? C_NULL : pointer(sol[1].entries)        # blocks are views of ONE entries vector (FEVector)
        GC.@preserve keep sol check(gs.ctx.ptr, ccall((:extfem_assemble_bilinear, lib), Cint,
            (Ptr{Cvoid}, Cint, Ref{OpDesc}, Ptr{Float64}, Cint, Ptr{Float64}), gs.ctx.ptr, gs.pattern, d, solv, 0, gs.stage_nz))
        A.cscmatrix.nzval .+= gs.stage_nz
        return nothing
    end
    return O
end

# LinearOperator: O.assembler(b.entries[, sol blocks]; time)   (linear_operator.jl:642, :710-713)
function gpu!(O::LinearOperator, gs::GPUSystem, SC; blocks_test, blocks_args = Int[])
    O.assembler = function (b, sol = nothing; time = 0.0, kwargs...)
        keep = Any[]
        kern = isempty(O.ops_args) ? O.kernel : ExtendableFEMBase.standard_kernel
        d = opdesc(O.parameters, kern, blocks_test, O.ops_test, Int[], DataType[], blocks_args, O.ops_args; time, keep)
        solv = sol === nothing ? C_NULL : pointer(sol[1].entries)
        GC.@preserve keep sol check(gs.ctx.ptr, ccall((:extfem_assemble_linear, lib), Cint,
            (Ptr{Cvoid}, Cint, Ref{OpDesc}, Ptr{Float64}, Cint, Ptr{Float64}), gs.ctx.ptr, gs.pattern, d, solv, 0, gs.stage_b))
        b .+= gs.stage_b
        return nothing
    end
    return O
end

# NonlinearOperator: O.assembler(A.entries, b.entries, sol blocks; time)   (nonlinear_operator.jl:440, :491); the library
# returns the Newton matrix and the right-hand side J u - F(u) of the linearisation in one call
function gpu!(O::NonlinearOperator, gs::GPUSystem, SC; blocks_test, blocks_args = blocks_test)
    O.assembler = function (A, b, sol; time = 0.0, kwargs...)
        keep = Any[]
        d = opdesc(O.parameters, O.kernel.kernel, blocks_test, O.ops_test, Int[], DataType[], blocks_args, O.ops_args; time, keep)
        solv = pointer(sol[1].entries)
        GC.@preserve keep sol check(gs.ctx.ptr, ccall((:extfem_assemble_nonlinear, lib), Cint,
            (Ptr{Cvoid}, Cint, Ref{OpDesc}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}),
            gs.ctx.ptr, gs.pattern, d, solv, 0, gs.stage_nz, gs.stage_b))
        A.cscmatrix.nzval .+= gs.stage_nz
        b .+= gs.stage_b
        return nothing
    end
    return O
end

# ItemIntegrator: evaluate(O, sol) (item_integrator.jl:323-352)
function gpu_evaluate(O::ItemIntegrator, gs::GPUSystem, sol; blocks_args, time = 0.0)
    keep = Any[]
    d = opdesc(O.parameters, O.kernel, Int[], DataType[], Int[], DataType[], blocks_args, O.ops_args; time, keep)
    ncells = num_cells(sol[1].FES.xgrid)
    rd = O.parameters[:resultdim] == 0 ? sum(Length4Operator(op, dim_grid(sol[1].FES.xgrid), get_ncomponents(eltype(sol[j].FES)))
                                            for (op, j) in zip(O.ops_args, blocks_args)) : O.parameters[:resultdim]
    out = O.parameters[:piecewise] ? zeros(rd, ncells) : zeros(rd)
    GC.@preserve keep sol check(gs.ctx.ptr, ccall((:extfem_integrate, lib), Cint,
        (Ptr{Cvoid}, Cint, Ref{OpDesc}, Ptr{Float64}, Cint, Cint, Ptr{Float64}),
        gs.ctx.ptr, gs.pattern, d, pointer(sol[1].entries), rd, O.parameters[:piecewise], out))
    return out
end

# ---- device-resident system (opt-in fast path: nothing but the solution crosses PCIe) ---------------------------------
# fill!(nzval, 0) / fill!(b, 0) of assemble_system! (solvers.jl:130-135)
values_zero!(gs; matrix = true, rhs = true) =
    check(gs.ctx.ptr, ccall((:extfem_values_zero, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cint), gs.ctx.ptr, gs.pattern, matrix, rhs))
# one operator into the device-resident system (accumulate = true: on top of what the previous operators left)
assemble_resident!(gs, d::OpDesc; sol = C_NULL, kind = :bilinear) =
    kind == :nonlinear ?
    check(gs.ctx.ptr, ccall((:extfem_assemble_nonlinear, lib), Cint, (Ptr{Cvoid}, Cint, Ref{OpDesc}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}),
                            gs.ctx.ptr, gs.pattern, d, sol, 1, C_NULL, C_NULL)) :
    check(gs.ctx.ptr, ccall((kind == :linear ? :extfem_assemble_linear : :extfem_assemble_bilinear, lib), Cint,
                            (Ptr{Cvoid}, Cint, Ref{OpDesc}, Ptr{Float64}, Cint, Ptr{Float64}), gs.ctx.ptr, gs.pattern, d, sol, 1, C_NULL))
# apply_penalties! (homogeneousdata_operator.jl:186-201, interpolateboundarydata_operator.jl:199-214) incl. the assemble_sol leg
function penalties!(gs, bdofs::Vector{Int}, values, penalty, sol::Vector{Float64})
    v = values === nothing ? C_NULL : pointer(values)
    GC.@preserve values check(gs.ctx.ptr, ccall((:extfem_apply_penalties, lib), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Float64), gs.ctx.ptr, gs.pattern, length(bdofs), bdofs, v, penalty))
    GC.@preserve values check(gs.ctx.ptr, ccall((:extfem_apply_values, lib), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Int64), gs.ctx.ptr, length(bdofs), bdofs, v, sol, length(sol)))
end
# compute_nonlinear_residual! (solvers.jl:38-43): residual = b - A * sol on the device-resident system
function residual!(res::Vector{Float64}, gs, sol::Vector{Float64})
    check(gs.ctx.ptr, ccall((:extfem_residual, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}), gs.ctx.ptr, gs.pattern, sol, res))
    return res
end
# download of the assembled system (when a host direct solver takes over)
function values_get!(A::FEMatrix, b::FEVector, gs)
    check(gs.ctx.ptr, ccall((:extfem_values_get, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}),
                            gs.ctx.ptr, gs.pattern, A.entries.cscmatrix.nzval, b.entries))
end
# symmetric forms: only the lower triangle crosses PCIe (half the bytes); the result is what CHOLMOD / cg take of an SPD matrix
function lower_system(gs, b::Vector{Float64})
    n = Ref{Int64}(0)
    sig = (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Int64})
    check(gs.ctx.ptr, ccall((:extfem_pattern_get_lower, lib), Cint, sig, gs.ctx.ptr, gs.pattern, n, C_NULL, C_NULL))
    colptr = Vector{Int64}(undef, gs.nrows + 1); rowval = Vector{Int64}(undef, n[]); nzval = Vector{Float64}(undef, n[])
    check(gs.ctx.ptr, ccall((:extfem_pattern_get_lower, lib), Cint, sig, gs.ctx.ptr, gs.pattern, C_NULL, colptr, rowval))
    check(gs.ctx.ptr, ccall((:extfem_values_get_lower, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}),
                            gs.ctx.ptr, gs.pattern, nzval, b))
    return Symmetric(SparseMatrixCSC(gs.nrows, gs.nrows, colptr, rowval, nzval), :L)
end
end # module
