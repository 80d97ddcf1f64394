"""FESpace / FEVector / FEMatrix stand-ins (H1 Lagrange P1/P2 on simplices).

The reference takes these from ExtendableFEMBase.jl (not in /root/reference).  What the
hot path consumes is only the cell dof map ``FES[CellDofs]`` (helper_functions.jl:561-567)
plus block offsets (``FE_test[j].offset``, bilinear_operator.jl:770-771), so that is what
this mirror provides.  Dof numbering (engine convention, matches H1P1/H1P2/H1Pk<=2 of
ExtendableFEMBase as far as recalled): per component [nodes..., edges...], components
stacked with stride ``nnodes (+ nedges)``.  Local order per component: vertices in
cell-node order, then edges in local-edge order (tri: 12,23,31; tet: 12,13,14,23,24,34).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .grids import ExtendableGrid

# element ids shared with include/extfem_cuda.h
EXTFEM_FE_H1P1 = 1
EXTFEM_FE_H1P2 = 2

__all__ = ["FESpace", "H1P1", "H1P2", "H1Pk", "FEVector", "FEVectorBlock", "FEMatrix",
           "EXTFEM_FE_H1P1", "EXTFEM_FE_H1P2", "interpolate"]


@dataclass(frozen=True)
class FEType:
    name: str
    ncomponents: int
    edim: int
    order: int

    @property
    def fe_id(self) -> int:
        return {1: EXTFEM_FE_H1P1, 2: EXTFEM_FE_H1P2}[self.order]


def H1P1(ncomponents: int, edim: int | None = None) -> FEType:
    return FEType("H1P1", ncomponents, edim or 0, 1)


def H1P2(ncomponents: int, edim: int) -> FEType:
    return FEType("H1P2", ncomponents, edim, 2)


def H1Pk(ncomponents: int, edim: int, order: int) -> FEType:
    if order not in (1, 2):
        raise NotImplementedError(
            "H1Pk with order > 2 is not supported by the B200 engine (EXTFEM_ERR_UNSUPPORTED_ELEMENT)")
    return FEType("H1Pk", ncomponents, edim, order)


class FESpace:
    """``FESpace{FEType}(xgrid)``."""

    def __init__(self, fetype: FEType, xgrid: ExtendableGrid):
        self.fetype = fetype
        self.xgrid = xgrid
        g = xgrid
        nn = g.nnodes
        if fetype.order == 1:
            scalar = g.cellnodes.astype(np.int64)
            nscalar = nn
            bscalar = g.bfacenodes.astype(np.int64)
        elif g.dim == 1:
            # 1D: the "edge" dof of P2 is the cell midpoint
            mid = nn + np.arange(1, g.ncells + 1, dtype=np.int64)
            scalar = np.concatenate([g.cellnodes.astype(np.int64), mid[:, None]], axis=1)
            nscalar = nn + g.ncells
            bscalar = g.bfacenodes.astype(np.int64)
        else:
            edgenodes, celledges = g.edges()
            scalar = np.concatenate([g.cellnodes.astype(np.int64), celledges.astype(np.int64) + nn], axis=1)
            nscalar = nn + edgenodes.shape[0]
            bscalar = self._bface_scalar_dofs(g, edgenodes, nn)
        nc = fetype.ncomponents
        self.coffset = nscalar
        self.ndofs = nc * nscalar
        self.nscalar_per_cell = scalar.shape[1]
        self.celldofs = np.concatenate([scalar + c * nscalar for c in range(nc)], axis=1).astype(np.int32)
        self.bfacedofs = np.concatenate([bscalar + c * nscalar for c in range(nc)], axis=1).astype(np.int32)

    @staticmethod
    def _bface_scalar_dofs(g, edgenodes, nn):
        bn = g.bfacenodes.astype(np.int64)
        if g.dim == 1:
            return bn
        en = edgenodes.astype(np.int64)
        keys = en[:, 0] * (nn + 1) + en[:, 1]
        loc = {2: [(0, 1)], 3: [(0, 1), (1, 2), (2, 0)]}[g.dim]
        cols = [bn]
        for a, b in loc:
            lo, hi = np.minimum(bn[:, a], bn[:, b]), np.maximum(bn[:, a], bn[:, b])
            cols.append((np.searchsorted(keys, lo * (nn + 1) + hi) + nn + 1)[:, None])
        return np.concatenate(cols, axis=1)

    @property
    def ndofs4cell(self) -> int:
        return self.celldofs.shape[1]

    def dof_coordinates(self) -> np.ndarray:
        """Coordinates of the scalar Lagrange points (nodes, then edge midpoints)."""
        g = self.xgrid
        if self.fetype.order == 1:
            return g.coords
        if g.dim == 1:
            cn = g.cellnodes.astype(np.int64)
            return np.concatenate([g.coords, 0.5 * (g.coords[cn[:, 0] - 1] + g.coords[cn[:, 1] - 1])])
        en = g.edges()[0].astype(np.int64)
        return np.concatenate([g.coords, 0.5 * (g.coords[en[:, 0] - 1] + g.coords[en[:, 1] - 1])])


class FEVectorBlock:
    def __init__(self, parent: "FEVector", FES: FESpace, offset: int):
        self.parent, self.FES, self.offset = parent, FES, offset

    @property
    def view(self) -> np.ndarray:
        return self.parent.entries[self.offset:self.offset + self.FES.ndofs]

    def __len__(self):
        return self.FES.ndofs


class FEVector:
    """``FEVector(FES)``: block vector with contiguous ``entries``."""

    def __init__(self, FES):
        FES = list(FES) if isinstance(FES, (list, tuple)) else [FES]
        self.FES = FES
        offs = np.concatenate([[0], np.cumsum([F.ndofs for F in FES])])
        self.entries = np.zeros(int(offs[-1]))
        self.blocks = [FEVectorBlock(self, F, int(o)) for F, o in zip(FES, offs[:-1])]

    def __getitem__(self, j) -> FEVectorBlock:
        return self.blocks[j]

    def __len__(self):
        return len(self.blocks)


class FEMatrix:
    """``FEMatrix(FES)``: block matrix; ``entries`` is one global CSC matrix
    (colptr/rowval/nzval, Int64 1-based like ``A.entries.cscmatrix``,
    src/solver_config.jl:190, src/solvers.jl:134)."""

    def __init__(self, FES_rows, FES_cols=None):
        FES_rows = list(FES_rows) if isinstance(FES_rows, (list, tuple)) else [FES_rows]
        FES_cols = FES_rows if FES_cols is None else (
            list(FES_cols) if isinstance(FES_cols, (list, tuple)) else [FES_cols])
        self.FES, self.FESY = FES_rows, FES_cols
        self.row_offsets = np.concatenate([[0], np.cumsum([F.ndofs for F in FES_rows])]).astype(np.int64)
        self.col_offsets = np.concatenate([[0], np.cumsum([F.ndofs for F in FES_cols])]).astype(np.int64)
        self.colptr = None
        self.rowval = None
        self.nzval = None
        self.pattern = None      # engine pattern handle wrapper

    @property
    def shape(self):
        return int(self.row_offsets[-1]), int(self.col_offsets[-1])

    def tocsc(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=self.shape)


def interpolate(block: FEVectorBlock, f) -> None:
    """Nodal Lagrange interpolation ``interpolate!(u[j], f)`` (exact for data in the space).
    ``f(x) -> array[ncomp]`` is called with x of shape [npoints, dim] and must vectorise."""
    F = block.FES
    pts = F.dof_coordinates()
    vals = np.asarray(f(pts), dtype=np.float64)
    if vals.ndim == 1:
        vals = vals[:, None]
    nc = F.fetype.ncomponents
    assert vals.shape == (pts.shape[0], nc), vals.shape
    v = block.view
    for c in range(nc):
        v[c * F.coffset:(c + 1) * F.coffset] = vals[:, c]
