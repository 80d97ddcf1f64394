"""Shared helpers of the parity tests: build the same system in the engine and in the oracle."""
import numpy as np

# tolerance stated by BASELINE.json north_star: values / rhs / residuals within 1e-12 relative
# (summation-order differences only).  Two forms are used:
#   check_values            normwise:  max|got - ref| <= 1e-12 * max|ref|
#   check_values_entrywise  entrywise, backward-error sense:  |got_i - ref_i| <= 1e-12 * sum_cells |contribution_i|,
#                           the scale coming from the oracle run with ora.abs_accumulate() -- an entry that is
#                           structurally present but cancels to ~0 cannot be compared relative to itself, and small
#                           entries (graded meshes, pressure couplings) are held to THEIR scale, not to the largest entry.
RTOL = 1e-12


class System:
    """Engine-side handles + oracle-side descriptions of one block system on one grid."""

    def __init__(self, pkg, ora, eng, grid, fetypes, block_coupling=None):
        self.pkg, self.ora, self.eng, self.grid = pkg, ora, eng, grid
        self.FES = [pkg.FESpace(t, grid) for t in fetypes]
        self.offsets = np.concatenate([[0], np.cumsum([F.ndofs for F in self.FES])]).astype(np.int64)
        self.N = int(self.offsets[-1])
        self.omesh = ora.Mesh(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes)
        if eng is not None:
            self.mesh = eng.mesh_set(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes)
            self.spaces = [eng.fespace_set(self.mesh, F) for F in self.FES]
            self.pat = eng.pattern_build(self.spaces, None, block_coupling)
            self.colptr, self.rowval = eng.pattern_get(self.pat)

    def oarg(self, block, op, offdiag=1.0):
        F = self.FES[block]
        return self.ora.OraArg(F.celldofs, F.fetype.ncomponents, F.fetype.order, op, int(self.offsets[block]), offdiag)

    def oargs(self, lst):
        return [self.oarg(b, o) for b, o in lst]


def check_values(got, ref, scale=None, rtol=RTOL, what="values"):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite entries"
    s = np.abs(ref).max() if scale is None else scale
    err = np.abs(got - ref).max()
    assert err <= rtol * s, f"{what}: max abs err {err:.3e} > {rtol:.0e} * scale {s:.3e}"
    return err / s if s > 0 else 0.0


def check_values_entrywise(got, ref, absscale, rtol=RTOL, what="values"):
    got = np.asarray(got); ref = np.asarray(ref); absscale = np.asarray(absscale)
    assert got.shape == ref.shape == absscale.shape, (got.shape, ref.shape, absscale.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite entries"
    err = np.abs(got - ref)
    # floor: an entry whose terms are all EXACTLY zero in the oracle (closed-form basis) may carry rounding of the inputs
    # (e.g. host-supplied polynomial coefficients), at the level of 1e-15 of the largest scale
    bad = err > rtol * (absscale + 1e-3 * absscale.max())
    assert not bad.any(), (f"{what}: {bad.sum()} entries beyond {rtol:.0e} * sum|contributions|; worst ratio "
                           f"{(err[bad] / np.maximum(absscale[bad], 1e-300)).max():.3e}")
    nzs = absscale > 0
    return float((err[nzs] / absscale[nzs]).max()) if nzs.any() else 0.0


def csc_subset(colptr_a, rowval_a, colptr_b, rowval_b):
    """True when pattern a is contained in pattern b (both 1-based CSC, sorted)."""
    ncols = colptr_a.size - 1
    cols_a = np.repeat(np.arange(ncols), np.diff(colptr_a))
    cols_b = np.repeat(np.arange(ncols), np.diff(colptr_b))
    nrows = int(max(rowval_a.max(), rowval_b.max())) + 1
    ka = cols_a.astype(np.int64) * nrows + rowval_a
    kb = cols_b.astype(np.int64) * nrows + rowval_b
    return np.isin(ka, kb).all()


class OracleBackend:
    """The CPU oracle behind the backend interface of host/problem.py (tests only): the same ProblemDescription that
    drives libextfem_cuda.so through EngineBackend runs on oracle/assembly_ref.c, into the same structural pattern."""

    def __init__(self, pkg, ora, FES):
        self.pkg, self.ora, self.FES = pkg, ora, FES
        g = FES[0].xgrid
        self.omesh = ora.Mesh(g.coords, g.cellnodes, g.cellregions, g.cellvolumes)
        self.bmesh = ora.Mesh(g.coords, g.bfacenodes, g.bfaceregions, g.bfacevolumes)
        self.offsets = np.concatenate([[0], np.cumsum([F.ndofs for F in FES])]).astype(np.int64)
        self.N = int(self.offsets[-1])
        allargs = [self._arg(j, 0) for j in range(len(FES))]
        self.colptr, self.rowval = ora.structural_pattern(allargs, allargs, (self.N, self.N))
        self.nz = np.zeros(self.rowval.size)
        self.b = np.zeros(self.N)

    def _arg(self, block, op, faces=False):
        F = self.FES[block]
        dofs = F.bfacedofs if faces else F.celldofs
        return self.ora.OraArg(dofs, F.fetype.ncomponents, F.fetype.order, op, int(self.offsets[block]))

    def zero(self):
        self.nz[:] = 0.0
        self.b[:] = 0.0

    def assemble(self, op, blocks, sol, time=0.0):
        P = op.parameters
        faces = P["entities"] == self.pkg.problem.ON_BFACES
        mesh = self.bmesh if faces else self.omesh
        mk = lambda oa: [self._arg(blocks[u], o, faces) for u, o in oa]     # noqa: E731
        kw = dict(params=P["params"], factor=P["factor"], quadorder=P["quadorder"], bonus_quadorder=P["bonus_quadorder"],
                  regions=list(P["regions"]), time=time)
        csc = (self.colptr, self.rowval)
        if op.kind == "bilinear":
            args = mk(op.oa_args)
            self.nz += self.ora.assemble_bilinear(mesh, mk(op.oa_test), mk(op.oa_ansatz), op.kernel, args=args, sol=sol if args else None,
                                                  args_sol_offsets=[a.offset for a in args], transposed_copy=P["transposed_copy"],
                                                  lump=P["lump"], csc=csc, **kw)
        elif op.kind == "linear":
            args = mk(op.oa_args)
            self.ora.assemble_linear(mesh, mk(op.oa_test), self.b, op.kernel, args=args, sol=sol if args else None,
                                     args_sol_offsets=[a.offset for a in args], tabulated=P.get("tabulated"), **kw)
        else:
            nz, _ = self.ora.assemble_nonlinear(mesh, mk(op.oa_test), mk(op.oa_args), sol, self.b, op.kernel, csc=csc, **kw)
            self.nz += nz

    def penalties(self, dofs, values, penalty):
        for d, v in zip(np.asarray(dofs) - 1, values):      # apply_penalties!: homogeneousdata_operator.jl:186-201
            lo, hi = self.colptr[d] - 1, self.colptr[d + 1] - 1
            k = lo + np.searchsorted(self.rowval[lo:hi], d + 1)
            assert self.rowval[k] == d + 1
            self.nz[k] = penalty
            self.b[d] = penalty * v

    def system(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nz.copy(), self.rowval - 1, self.colptr - 1), shape=(self.N, self.N)), self.b.copy()

    def residual(self, sol):
        A, b = self.system()
        return b - A @ sol

    def integrate(self, op, blocks, sol, resultdim, time=0.0):
        P = op.parameters
        args = [self._arg(blocks[u], o) for u, o in op.oa_args]
        return self.ora.integrate(self.omesh, args, sol, op.kernel, params=P["params"], factor=P["factor"], quadorder=P["quadorder"],
                                  bonus_quadorder=P["bonus_quadorder"], regions=list(P["regions"]), resultdim=resultdim, time=time,
                                  tabulated=P.get("tabulated"), piecewise=P["piecewise"])
