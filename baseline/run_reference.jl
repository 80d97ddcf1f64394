# baseline/run_reference.jl -- dumps what the UNMODIFIED reference assembles, for a direct comparison with libextfem_cuda.so.
#
# NOT EXECUTED in this repository (no Julia in the build image; scripts/probe_reference.sh looks for one on the GPU box and runs
# this file when it finds it).  Usage:  julia --project=<ExtendableFEM.jl checkout> baseline/run_reference.jl <outdir> [n2d] [n3d]
#
# For each case it writes <outdir>/<case>.bin:
#   int64 header {dim, ncells, nnodes, nd, ndofs, nnz}; coords f64[dim*nnodes]; cellnodes i64[(dim+1)*ncells]; celldofs i64[nd*ncells];
#   colptr i64[ndofs+1]; rowval i64[nnz]; nzval f64[nnz]; b f64[ndofs]
# (the layout tests/abi_driver.c reads).  tests/test_reference_dump.py feeds mesh and dofmap to the engine and compares
# pattern (the reference's is a subset of the structural one), values and rhs to 1e-12.
using ExtendableFEM, ExtendableFEMBase, ExtendableGrids, SparseArrays

f201!(result, qpinfo) = (result[1] = qpinfo.x[1] * qpinfo.x[2]; nothing)   # Example201:37-40 == registry kernel "xy" (any dim)

function dump_case(path, xgrid, FEType, f!)
    FES = FESpace{FEType}(xgrid)
    A = FEMatrix(FES)
    b = FEVector(FES)
    assemble!(A, BilinearOperator([grad(1)]; factor = 1.0))               # standard_kernel
    assemble!(b, LinearOperator(f!, [id(1)]))
    flush!(A.entries)
    S = A.entries.cscmatrix
    coords = xgrid[Coordinates]
    cn = xgrid[CellNodes]
    cd = FES[CellDofs]
    dim, nnodes = size(coords)
    ncells = num_cells(xgrid)
    nd = max_num_targets_per_source(cd)
    open(path, "w") do io
        write(io, Int64[dim, ncells, nnodes, nd, FES.ndofs, nnz(S)])
        write(io, Float64.(coords))
        write(io, Int64[cn[j, c] for j in 1:(dim + 1), c in 1:ncells])
        write(io, Int64[cd[j, c] for j in 1:nd, c in 1:ncells])
        write(io, Int64.(S.colptr)); write(io, Int64.(S.rowval)); write(io, Float64.(S.nzval)); write(io, Float64.(b.entries))
    end
    @info "wrote $path" ncells FES.ndofs nnz(S)
end

function main(outdir, n2d = 16, n3d = 8)
    mkpath(outdir)
    X2 = range(0, 1; length = n2d + 1)
    X3 = range(0, 1; length = n3d + 1)
    dump_case(joinpath(outdir, "config1_p3_2d.bin"), simplexgrid(X2, X2), H1Pk{1, 2, 3}, f201!)       # README.md:50 / Example201:66
    dump_case(joinpath(outdir, "config2_p2_3d.bin"), simplexgrid(X3, X3, X3), H1P2{1, 3}, f201!)      # Example301:60-80
    dump_case(joinpath(outdir, "p2_2d.bin"), simplexgrid(X2, X2), H1P2{1, 2}, f201!)
    dump_case(joinpath(outdir, "p1_3d.bin"), simplexgrid(X3, X3, X3), H1P1{1}, f201!)
end

main(ARGS[1], (length(ARGS) > 1 ? parse(Int, ARGS[2]) : 16), (length(ARGS) > 2 ? parse(Int, ARGS[3]) : 8))
