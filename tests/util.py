"""Shared helpers of the parity tests: build the same system in the engine and in the oracle."""
import numpy as np

# tolerance stated by BASELINE.json north_star: values / rhs / residuals within 1e-12 relative
# (summation-order differences only).  "Relative" is taken entrywise against the magnitude of
# the sum of absolute cell contributions (backward-error sense), bounded below by normwise scale:
# an entry that is structurally present but cancels to ~0 cannot be compared relative to itself.
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
            self.spaces = [eng.space_set(self.mesh, F.fetype.fe_id, F.fetype.ncomponents, F.celldofs, F.ndofs)
                           for F in self.FES]
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


def csc_subset(colptr_a, rowval_a, colptr_b, rowval_b):
    """True when pattern a is contained in pattern b (both 1-based CSC, sorted)."""
    ncols = colptr_a.size - 1
    cols_a = np.repeat(np.arange(ncols), np.diff(colptr_a))
    cols_b = np.repeat(np.arange(ncols), np.diff(colptr_b))
    nrows = int(max(rowval_a.max(), rowval_b.max())) + 1
    ka = cols_a.astype(np.int64) * nrows + rowval_a
    kb = cols_b.astype(np.int64) * nrows + rowval_b
    return np.isin(ka, kb).all()
