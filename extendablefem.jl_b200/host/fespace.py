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
EXTFEM_FE_TABULATED = 100

__all__ = ["FESpace", "H1P1", "H1P2", "H1Pk", "FEVector", "FEVectorBlock", "FEMatrix",
           "EXTFEM_FE_H1P1", "EXTFEM_FE_H1P2", "EXTFEM_FE_TABULATED", "interpolate", "monomial_exponents",
           "lagrange_nodes_p3", "nodal_basis_coefficients"]


@dataclass(frozen=True)
class FEType:
    name: str
    ncomponents: int
    edim: int
    order: int

    @property
    def fe_id(self) -> int:
        # order 3 is not built into the engine: it goes in as a host-supplied polynomial basis (extfem_space_set_tables)
        return {1: EXTFEM_FE_H1P1, 2: EXTFEM_FE_H1P2, 3: EXTFEM_FE_TABULATED}[self.order]


def H1P1(ncomponents: int, edim: int | None = None) -> FEType:
    return FEType("H1P1", ncomponents, edim or 0, 1)


def H1P2(ncomponents: int, edim: int) -> FEType:
    return FEType("H1P2", ncomponents, edim, 2)


def H1Pk(ncomponents: int, edim: int, order: int) -> FEType:
    """``H1Pk{ncomponents, edim, order}``.  Orders 1 and 2 are built into the engine; order 3 (README.md:50, Example201:66;
    edim <= 2 here) is handed to it as a polynomial reference basis."""
    if order not in (1, 2, 3) or (order == 3 and edim > 2):
        raise NotImplementedError("H1Pk: orders 1-2 (any dimension) and order 3 (edim <= 2) are provided")
    return FEType("H1Pk", ncomponents, edim, order)


def monomial_exponents(order: int, dim: int):
    """Monomial enumeration of extfem_space_set_tables: x^i y^j z^k, ``for k: for j: for i`` (i fastest)."""
    out = []
    for k in range(order + 1 if dim >= 3 else 1):
        for j in range(order - k + 1 if dim >= 2 else 1):
            for i in range(order - k - j + 1):
                out.append((i, j, k)[:max(dim, 1)])
    return out


def lagrange_nodes_p3(dim: int) -> np.ndarray:
    """Reference Lagrange points of the cubic element in local dof order: vertices, two per edge (a,b) at 1/3 and 2/3 from a
    to b (local edge order of the grid), then the cell centre (2D)."""
    if dim == 0:
        return np.zeros((1, 0))
    verts = np.concatenate([np.zeros((1, dim)), np.eye(dim)])
    edges = {1: [(0, 1)], 2: [(0, 1), (1, 2), (2, 0)]}[dim]
    pts = [v for v in verts]
    for a, b in edges:
        pts.append(verts[a] + (verts[b] - verts[a]) / 3.0)
        pts.append(verts[a] + 2.0 * (verts[b] - verts[a]) / 3.0)
    if dim == 2:
        pts.append(verts.mean(axis=0))
    return np.array(pts)


def nodal_basis_coefficients(nodes: np.ndarray, order: int) -> np.ndarray:
    """coeffs[j][m]: basis_j = sum_m coeffs[j][m] * monomial_m with basis_j(node_i) = delta_ij."""
    dim = nodes.shape[1]
    if dim == 0:
        return np.ones((1, 1))
    ex = monomial_exponents(order, dim)
    V = np.stack([np.prod(nodes ** np.array(e)[None, :], axis=1) for e in ex], axis=1)   # [node][monomial]
    assert V.shape[0] == V.shape[1], V.shape
    return np.ascontiguousarray(np.linalg.inv(V).T)


class FESpace:
    """``FESpace{FEType}(xgrid)``."""

    def __init__(self, fetype: FEType, xgrid: ExtendableGrid):
        self.fetype = fetype
        self.xgrid = xgrid
        g = xgrid
        nn = g.nnodes
        self.ref_coeffs = self.ref_coeffs_bface = None
        if fetype.order == 3:
            scalar, bscalar, nscalar = self._p3_dofs(g)
            self.ref_coeffs = nodal_basis_coefficients(lagrange_nodes_p3(g.dim), 3)
            self.ref_coeffs_bface = nodal_basis_coefficients(lagrange_nodes_p3(g.dim - 1), 3)
        elif fetype.order == 1:
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
    def _p3_dofs(g):
        """Cubic Lagrange dofs: nodes, then two per edge -- the first one next to the edge's LOWER global node --, then one
        per cell (2D).  A cell lists the two dofs of a local edge (a,b) in the order (next to a, next to b), so the reference
        basis is the same in every cell and the edge orientation lives in the dof map."""
        nn = g.nnodes
        cn = g.cellnodes.astype(np.int64)
        bn = g.bfacenodes.astype(np.int64)
        if g.dim == 1:
            first = nn + 2 * np.arange(g.ncells, dtype=np.int64) + 1
            flip = cn[:, 0] > cn[:, 1]
            e1 = np.where(flip, first + 1, first); e2 = np.where(flip, first, first + 1)
            return np.concatenate([cn, e1[:, None], e2[:, None]], axis=1), bn, nn + 2 * g.ncells
        edgenodes, celledges = g.edges()
        ne = edgenodes.shape[0]
        loc = {2: [(0, 1), (1, 2), (2, 0)]}[g.dim]
        cols = [cn]
        for le, (a, b) in enumerate(loc):
            first = nn + 2 * (celledges[:, le].astype(np.int64) - 1) + 1
            flip = cn[:, a] > cn[:, b]
            cols += [np.where(flip, first + 1, first)[:, None], np.where(flip, first, first + 1)[:, None]]
        cols.append((nn + 2 * ne + np.arange(1, g.ncells + 1, dtype=np.int64))[:, None])
        # boundary faces (edges): nodes, then the edge's two dofs in the face's own direction
        en = edgenodes.astype(np.int64)
        keys = en[:, 0] * (nn + 1) + en[:, 1]
        lo, hi = np.minimum(bn[:, 0], bn[:, 1]), np.maximum(bn[:, 0], bn[:, 1])
        first = nn + 2 * np.searchsorted(keys, lo * (nn + 1) + hi) + 1
        flip = bn[:, 0] > bn[:, 1]
        bcols = [bn, np.where(flip, first + 1, first)[:, None], np.where(flip, first, first + 1)[:, None]]
        return np.concatenate(cols, axis=1), np.concatenate(bcols, axis=1), nn + 2 * ne + g.ncells

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
        if self.fetype.order == 3:
            # every dof is a Lagrange point: read it off the cells (reference points in local dof order)
            ref = lagrange_nodes_p3(g.dim)
            lam = np.concatenate([1.0 - ref.sum(axis=1, keepdims=True), ref], axis=1)            # [nloc, dim+1]
            x = np.einsum("lv,cvd->cld", lam, g.coords[g.cellnodes.astype(np.int64) - 1])         # [ncells, nloc, dim]
            out = np.zeros((self.coffset, g.dim))
            out[self.celldofs[:, :ref.shape[0]].astype(np.int64) - 1] = x
            return out
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
